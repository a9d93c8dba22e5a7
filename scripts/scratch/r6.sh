python -m pytest tests/test_gpu_index_build.py -x -q -m gpu 2>&1 | grep -E "^E|Error" | head -8
for T in 2048 4096 8192; do echo "tile $T"; ORAG_FP_TILE_DOCS=$T python scripts/debug_bm25.py 1250000 256 2048 --no-dense 2>&1 | grep -E "sparse iter 2|ms ==" | tail -2; done
