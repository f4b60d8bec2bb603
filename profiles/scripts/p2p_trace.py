"""Where the fused multi-GPU exchange spends its time: %globaltimer stamps of the tick kernel's last CTA on every rank
(mppi_comm_p2p_trace): words stored, every rank's words seen, nominal updated.  Launch like bench.py:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/scripts/p2p_trace.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-mppi-mpc_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from bench import K_PER_GPU, T_H, diffdrive_kwargs  # noqa: E402
from mppi_b200.mppi_differential_drive import MPPIAlgorithms  # noqa: E402

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctrl = MPPIAlgorithms(**diffdrive_kwargs(K_PER_GPU * world, T_H, 10.0), seed=7, device=local, rank=rank, world=world)
ctrl.comm_init_from_torch()
eng = ctrl.engine
st = torch.cuda.Stream()
eng.set_stream(st.cuda_stream)
eng.set_trace(True)
x0 = np.zeros(3)
tick = 0
rows = []
for rep in range(25):
    dist.barrier()
    for _ in range(20):                                   # free-running ticks: the GPUs pace each other through the exchange
        eng.step_async(x0, None, 7, tick); tick += 1
    eng.synchronize()
    t = eng.comm_p2p_trace()
    ctas, last = eng.trace()
    rows.append((t[1] - t[0], t[2] - t[1], t[3] - t[2], t[0] - ctas[:, 1].max(), t[3] - ctas[:, 0].min()))
r = torch.tensor(np.median(np.array(rows[3:], dtype=np.float64), axis=0) / 1e3, device="cuda")
allr = [torch.zeros_like(r) for _ in range(world)]
dist.all_gather(allr, r)
if rank == 0:
    a = torch.stack(allr).cpu().numpy()
    print("fused exchange, %d GPUs, K = %d per GPU, H = %d; medians over 22 free-running batches, microseconds per rank:" % (world, K_PER_GPU, T_H))
    print("%-58s %s" % ("", " ".join("r%-6d" % i for i in range(world))))
    for j, name in enumerate(("last rollouts done -> local partials merged", "words stored to every peer (no fence)",
                              "wait until every rank's words carry this tick's number", "rank-order merge + filter + update",
                              "first CTA start -> nominal updated (whole tick)")):
        col = {0: 3, 1: 0, 2: 1, 3: 2, 4: 4}[j]
        print("%-58s %s" % (name, " ".join("%7.2f" % v for v in a[:, col])))
dist.destroy_process_group()
