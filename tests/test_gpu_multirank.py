"""GPU, two ranks over NCCL (skipped with fewer than 2 GPUs): time-slice sharding of the fine solves with one all-gather,
the replicated sweep and the dimension-sharded sweep all give the bits of the single-GPU run, on every rank."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
rank, world, port = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
import numpy as np, torch, torch.distributed as dist
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib
from helpers import load_run, case_system, device_system
torch.cuda.set_device(rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
solo = [dist.new_group([r]) for r in range(world)][rank]      # a one-rank group: the single-GPU run inside this process
ok = True
for name in ("fhn_d32_N32_m12", "lorenz_N32_m11"):
    z, cfg, mkw = load_run(name)
    key, kw = case_system(name)
    outs = {}
    for label, group, shard in (("single", solo, False), ("replicated", None, False), ("sharded", None, True)):
        ode = device_system(key, **kw)
        solver = nn.CudaSolverRK(ode.get_vector_field(), handle=_lib.default_handle(rank), **{k: cfg[k] for k in ("Ng", "Nf", "F", "G")})
        p = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=cfg["N"], epsilon=float(z["epsilon"]), verbose='',
                              group=group, shard_sweep=shard)
        outs[label] = p.run(model='nngp', **mkw)
    for label in ("replicated", "sharded"):
        same = (outs[label]['k'] == outs["single"]['k'] and outs[label]['conv_int'] == outs["single"]['conv_int']
                and np.array_equal(outs[label]['u'], outs["single"]['u'])
                and np.array_equal(outs[label]['err'], outs["single"]['err'], equal_nan=True))
        print(f"rank {rank} {name} {label}: K={outs[label]['k']} bitwise equal to the single-GPU run: {same}", flush=True)
        ok &= same
    v = torch.from_numpy(outs["sharded"]['u']).cuda()
    lst = [torch.empty_like(v) for _ in range(world)]
    dist.all_gather(lst, v)
    ok &= all(torch.equal(lst[0], x) for x in lst)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


def test_two_rank_nccl_runs_equal_single_gpu_bitwise(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker_nccl.py"
    script.write_text(_WORKER)
    port = str(33500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", port], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for o in outs:
        print(o[-3000:])
    assert [p.returncode for p in procs] == [0, 0]
