set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "stencil or fused_modes or pattern" > gpurun_out/pytest_st2.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/pytest_st2.log
timeout 300 python tools/stencil_bench.py 512 > gpurun_out/stencil_bench_512_v2.log 2>&1; echo "bench rc=$?"
tail -40 gpurun_out/stencil_bench_512_v2.log
cat > /tmp/one.py <<'P'
import sys; sys.path.insert(0,'/root/repo')
import torch
from krylov_b200.generate import device_stencil7
from krylov_b200.device import Ops
A = device_stencil7(512,512,512); n=A.shape[0]; ops=Ops(n,1)
x=torch.randn(n,1,dtype=torch.float64,device='cuda'); y=torch.empty_like(x); out=ops.slots(1)
for _ in range(4): ops.spmv(A,x,y,dot=1,w=x,out=out[0])
torch.cuda.synchronize()
P
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_spmv_stencil2 -s 2 -c 1 -f -o gpurun_out/prof_stencil2_512 python /tmp/one.py > gpurun_out/ncu_st2.log 2>&1; echo "ncu rc=$?"
