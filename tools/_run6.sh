cd /root/repo
timeout 600 python bench.py > gpurun_out/bench_n1_final2.json 2> gpurun_out/bench_n1_final2.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/bench_n1_final2.json; tail -2 gpurun_out/bench_n1_final2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_n1_final2.json').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"phases",d["roofline"]["phases_ms"],"e2e",d["e2e"]["value"],d["e2e"]["solve_seconds"],"clocks",d["clocks"])
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_march.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_stencil_march -s 8 -c 2 -f -o gpurun_out/prof_march_cg_512_final python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 500 python tools/bench_configs.py > gpurun_out/configs_march.txt 2>&1; echo "configs rc=$?"
tail -30 gpurun_out/configs_march.txt
