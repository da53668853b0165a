import sys, time, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scikit-motionplan_b200')
import torch, numpy as np
import bench
from skmp_b200 import _lib
from skmp_b200.constraint import CollFreeConst, PairWiseSelfCollFreeConst
from skmp_b200.robot.jaxon import JaxonConfig
from skmp_b200.robot_model import Jaxon
pr2, colkin, grid = bench.make_problem()
const = CollFreeConst(colkin, grid, pr2)
L=_lib.lib()
def run(n, label, fn):
    for _ in range(50): fn()
    torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(2000): fn()
    t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    print(label, "N=%d"%n, "host %.1f us/call, incl sync %.1f us/call"%((t1-t)/2000*1e6,(t2-t)/2000*1e6))
for n in (1, 64):
    qs=torch.zeros((n,14),device='cuda'); vals=torch.empty((n,152),device='cuda'); jac=torch.empty((n,152,14),device='cuda'); valid=torch.empty((n,),dtype=torch.uint8,device='cuda')
    plan=const._plan(); st=C.c_void_p(torch.cuda.current_stream().cuda_stream)
    run(n,"C-ABI collfree all+jac", lambda: L.skb_collfree(plan.handle, grid.native(), _lib.F32, C.c_void_p(qs.data_ptr()), n, 0.0, _lib.MODE_ALL, C.c_void_p(vals.data_ptr()), C.c_void_p(jac.data_ptr()), None, st))
    run(n,"C-ABI collfree validity", lambda: L.skb_collfree(plan.handle, grid.native(), _lib.F32, C.c_void_p(qs.data_ptr()), n, 0.0, _lib.MODE_ALL, None, None, C.c_void_p(valid.data_ptr()), st))
    run(n,"python evaluate", lambda: const.evaluate(qs, True))
    run(n,"python is_valid_batch", lambda: const.is_valid_batch(qs))
