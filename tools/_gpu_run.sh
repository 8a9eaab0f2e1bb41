cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-configs --no-general --no-parity"
timeout 600 $B > gpurun_out/r2t_bench_short.json 2> gpurun_out/r2t_bench_short.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches_bench_n1.csv $B > gpurun_out/r2t_ncu1.log 2>&1
tools/ncu_summ.sh gpurun_out/r2t_march_cg kb_stencil_march 8 2 -- $B
python tools/spmv_general_bench.py --one banded_100 merge 20 6 > /dev/null 2>&1 && tools/ncu_summ.sh gpurun_out/r2t_merge_banded kb_spmv_merge 4 1 -- python tools/spmv_general_bench.py --one banded_100 merge 20 6
tools/ncu_summ.sh gpurun_out/r2t_merge_powerlaw "kb_spmv_merge|kb_merge_fix" 8 2 -- python tools/spmv_general_bench.py --one powerlaw merge 20 6
tools/ncu_summ.sh gpurun_out/r2t_small kb_cg_small 2 1 -- python tools/small_cg_bench.py 256
tools/ncu_summ.sh gpurun_out/r2t_lines_var kb_spmm_lines 6 1 -- python tools/bench_spmm_var.py --n 192
ls -la gpurun_out | head -30
