python -m pytest tests/test_gpu_index_build.py tests/test_gpu_parity.py tests/test_zz_gpu_parity_scale.py -x -q -m gpu 2>&1 | tail -3
python scripts/debug_bm25.py 10000000 256 2048 --no-dense 2>&1 | grep -E "index build|sparse iter|ms ==" | tail -4
python scripts/debug_bm25.py 1250000 256 2048 --no-dense 2>&1 | grep -E "sparse iter|ms ==" | tail -2
python scripts/debug_bm25.py 10000000 1024 2048 --no-dense 2>&1 | grep -E "sparse iter|ms ==" | tail -2
ORAG_BM25_WARM_START=0 python scripts/debug_bm25.py 10000000 256 2048 --no-dense 2>&1 | grep -E "sparse iter|ms ==" | tail -2
