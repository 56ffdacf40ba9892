import os, sys
sys.path.insert(0,'.')
import torch.multiprocessing as mp
from tests.test_dist_cpu import _worker, _free_port
if __name__ == "__main__":
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port=_free_port()
    ps=[ctx.Process(target=_worker,args=(r,2,port,q)) for r in range(2)]
    [p.start() for p in ps]
    [p.join(90) for p in ps]
    print([p.exitcode for p in ps])
    while not q.empty(): print(q.get())
