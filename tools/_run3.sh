cd /root/repo
cat > /tmp/cgsteps.py <<'P'
import sys; sys.path.insert(0,'/root/repo')
import torch
from krylov_b200.cg import FusedCG
from krylov_b200.generate import device_stencil7
A = device_stencil7(512,512,512); n=A.shape[0]
g = torch.Generator(device="cuda").manual_seed(0)
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
st = FusedCG(A, b.reshape(n, 1), torch.zeros(n, 1, dtype=torch.float64, device="cuda"), 0.0, 0.0)
st.run(6); torch.cuda.synchronize()
P
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_stencil_march -s 4 -c 2 -f -o gpurun_out/prof_march_cg_512 python /tmp/cgsteps.py > gpurun_out/ncu_march.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/ncu_march.log
