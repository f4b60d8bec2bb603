"""Sample sharding over 2 GPUs: the exchange of (min, sum w, sum w*eps) fused into the tick kernel over NVLink peer
memory (default) or as an NCCL all-gather; the sharded controller must produce the same nominal as the single-GPU one
(same global Philox samples), bit-identical on every rank, and a rank that never arrives must fail THAT tick loudly
without touching the nominal."""
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


def _racecar_worker(rank, world, port, K, T, ticks, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "dnn-mppi-mpc_b200"), os.path.join(root, "tests")]
    import torch.distributed as dist
    from mppi_b200.mppi_race_car_obstacle import MPPIRacecarController
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rc = MPPIRacecarController(horizon_step_T=T, number_of_samples_K=K, visualize_optimal_traj=False, visualze_sampled_trajs=False,
                               seed=13, device=rank, rank=rank, world=world)
    lp = rc.generate_lemniscate_trajectory(100, 10.0).astype(np.float32)
    rc.ref_path = lp
    rc.comm_init_from_torch()
    res = []
    for i in range(ticks):
        u0, u, _, _ = rc._calc_control_input(lp[i])
        res.append(u.copy())
    out.put((rank, np.array(res)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_racecar_small_K_time_parallel_rollout_with_fused_exchange():
    """configs[1] sharded: K = 16 384 race-car samples over 2 GPUs -- each rank's 8 192 go through the time-parallel rollout
    (CTAs of <= 64 samples) and the per-GPU triples through the exchange fused into the same kernel."""
    import torch.multiprocessing as mp
    from mppi_b200.mppi_race_car_obstacle import MPPIRacecarController
    K, T, ticks = 16384, 50, 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_racecar_worker, args=(r, 2, port, K, T, ticks, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(got[0], got[1])
    single = MPPIRacecarController(horizon_step_T=T, number_of_samples_K=K, visualize_optimal_traj=False, visualze_sampled_trajs=False, seed=13)
    lp = single.generate_lemniscate_trajectory(100, 10.0).astype(np.float32)
    single.ref_path = lp
    for i in range(ticks):
        u0, u, _, _ = single._calc_control_input(lp[i])
        assert np.max(np.abs(u - got[0][i])) <= 2e-5, (i, np.max(np.abs(u - got[0][i])))


def _timeout_worker(rank, world, port, out):
    import sys
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "dnn-mppi-mpc_b200"), os.path.join(root, "tests")]
    os.environ["MPPI_P2P_TIMEOUT_MS"] = "300"
    import torch.distributed as dist
    from mppi_b200 import MppiError
    from mppi_b200.mppi_differential_drive import MPPIAlgorithms
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctrl = MPPIAlgorithms(**_kwargs(1 << 14, 30), device=rank, rank=rank, world=world)
    ctrl.comm_init_from_torch()
    x = np.array([0.1, 0.05, 0.2])
    res = {}
    u0, u, _, _ = ctrl._calc_input_control(x)                 # tick 0: both ranks on time
    res["tick0"] = u.copy()
    dist.barrier()
    before = ctrl.engine.get_nominal().copy()
    idx_before = ctrl.prev_way_point_idx
    if rank == 1:
        time.sleep(1.5)                                       # rank 1 arrives 1.5 s late: rank 0's 300 ms guard expires
    try:
        ctrl._calc_input_control(x)
        res["tick1"] = "ok"
    except MppiError as e:
        res["tick1"] = str(e)
    res["nominal_untouched"] = bool(np.array_equal(ctrl.engine.get_nominal(), before)) and ctrl.prev_way_point_idx == idx_before
    dist.barrier()
    # the application resynchronises (here: both ranks reset the nominal) and carries on; the error does not stick
    ctrl.u_prev = np.zeros((30, 2))
    ctrl.prev_way_point_idx = 0
    dist.barrier()
    u0, u, _, _ = ctrl._calc_input_control(x)
    res["tick2"] = u.copy()
    out.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_exchange_peer_timeout_fails_the_tick_and_clears():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_timeout_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(got[0]["tick0"], got[1]["tick0"])
    assert "timed out" in got[0]["tick1"] and got[0]["nominal_untouched"], got[0]["tick1"]     # MPPI_E_NCCL, tick not applied
    assert got[1]["tick1"] == "ok"                           # the late rank found rank 0's words waiting
    assert np.array_equal(got[0]["tick2"], got[1]["tick2"])   # and the next tick is healthy on both
