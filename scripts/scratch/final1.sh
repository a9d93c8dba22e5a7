python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_gpu_tests_d.txt; tail -2 gpurun_out/r02_gpu_tests_d.txt
python bench.py > gpurun_out/r02_bench_config3_e.json 2> gpurun_out/b3.err
python bench.py --config 4 > gpurun_out/r02_bench_config4_d.json 2> gpurun_out/b4.err
python bench.py --config 2 > gpurun_out/r02_bench_config2_b.json 2> gpurun_out/b2.err
python bench.py --config pairwise > gpurun_out/r02_bench_pairwise_b.json 2> gpurun_out/bp.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_b.json 2> gpurun_out/br.err
for f in config3_e config4_d config2_b pairwise_b reference_b; do python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_$f.json").read().strip().splitlines()[-1]); r=d.get("roofline") or {}
print("$f", round(d["value"],1), d.get("ms_per_step"), "e2e", round(d["e2e"]["value"],1), "roof", r.get("launch_ms"), r.get("launch_ms_min"), r.get("frac"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("verified_against_oracle") or {}).get("bitwise"))
PY
done
python scripts/debug_bm25.py 10000000 256 2048 --no-dense > gpurun_out/dbg_d.txt 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bm25_ms_kernel -s 1 -c 1 -o gpurun_out/bm25_ms_d -f python scripts/debug_bm25.py 10000000 256 2048 --no-dense > gpurun_out/ncu_d.log 2>&1; tail -2 gpurun_out/ncu_d.log; grep "sparse iter 2" gpurun_out/dbg_d.txt
