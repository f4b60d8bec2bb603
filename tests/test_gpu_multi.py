"""Sample sharding over 2 GPUs with the NCCL all-gather of (min, sum w, sum w*eps): the sharded
controller must produce the same nominal as the single-GPU one (same global Philox samples)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _kwargs(K, T):
    from golden_util import Golden
    g = Golden("diffdrive_pe0.05")
    return dict(delta_t=0.1, ref_path=g.path, max_speed=5.0, max_omega=3.14, num_samples_K=K, num_horizons_T=T,
                param_exploration=0.05, param_lambda=1.0, param_alpha=0.2, sigma=np.array([[0.1, 0.0], [0.0, 0.01]]),
                stage_cost_weight=np.array([5.0, 5.0, 10.0]), terminal_cost_weight=np.array([5.0, 5.0, 10.0]),
                visualize_optimal_traj=False, visualze_sampled_trajs=False, cost_mode="sum", waypoint_mode="frozen",
                temperature=2.0, seed=21)


def _worker(rank, world, port, K, T, ticks, out, exchange):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "dnn-mppi-mpc_b200"), os.path.join(root, "tests")]
    import torch.distributed as dist
    from mppi_b200.mppi_differential_drive import MPPIAlgorithms
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctrl = MPPIAlgorithms(**_kwargs(K, T), device=rank, rank=rank, world=world)
    ctrl.comm_init_from_torch(exchange=exchange)
    x = np.array([0.1, 0.05, 0.2])
    res = []
    for _ in range(ticks):
        u0, u, _, _ = ctrl._calc_input_control(x)
        res.append(u.copy())
    out.put((rank, np.array(res)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpu_sharded_tick_equals_single_gpu(exchange):
    import torch.multiprocessing as mp
    from mppi_b200.mppi_differential_drive import MPPIAlgorithms
    K, T, ticks = 1 << 16, 50, 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, K, T, ticks, out, exchange)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(got[0], got[1])                      # bit-identical merge on every rank
    single = MPPIAlgorithms(**_kwargs(K, T))
    x = np.array([0.1, 0.05, 0.2])
    for i in range(ticks):
        u0, u, _, _ = single._calc_input_control(x)
        assert np.max(np.abs(u - got[0][i])) <= 2e-5, (i, np.max(np.abs(u - got[0][i])))
