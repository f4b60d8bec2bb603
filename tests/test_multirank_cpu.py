"""World-size-2 (gloo, CPU) coverage of the sample-sharded path's host logic: each rank evaluates its
contiguous shard of the samples with the oracle, publishes (rho, eta, sum w*eps), the ranks all-gather
the triples and merge them with the log-sum-exp rule of SURVEY.md 8e -- the rule the CUDA merge kernel
implements -- and must reproduce the unsharded tick.  Also exercises the NCCL-id hand-off helper."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from golden_util import Golden  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import mppi_oracle as orc  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def merge_triples(triples, temperature):
    """triples: (G, 2 + T*2) rows [rho_g, eta_g, N_g...] with N_g relative to rho_g."""
    rho = triples[:, 0].min()
    scale = np.exp(-(triples[:, 0] - rho) / temperature)
    eta = float((scale * triples[:, 1]).sum())
    N = (scale[:, None] * triples[:, 2:]).sum(axis=0)
    return rho, eta, N / eta


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = Golden("diffdrive_pe0.05")
    K, T = 512, 30
    sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    sp.temperature = 1.5
    x0, U = np.array([0.3, 0.2, 0.4]), np.zeros((T, 2))
    Kl = K // world
    shard = orc.diffdrive_spec(K=Kl, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    shard.temperature = 1.5
    # the explore/exploit split is by GLOBAL sample index (Q6): give the shard the global threshold
    S, _, _ = co.costs(shard, g.path, U, 0, x0, None, seed=9, tick=2, k_offset=rank * Kl, n_exploit=sp.n_exploit())
    eps = orc.philox_noise(9, 2, Kl, T, sp.sigma, k_offset=rank * Kl)
    rho = S.min()
    w = np.exp(-(S - rho) / sp.temperature)
    tri = np.concatenate([[rho, w.sum()], np.einsum("k,ktu->tu", w, eps).reshape(-1)])
    gathered = [torch.zeros(tri.size, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(tri))
    rho_m, eta_m, w_eps = merge_triples(torch.stack(gathered).numpy(), sp.temperature)
    # hand-off of an opaque 128-byte id, as comm_init_from_torch does
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    if rank == 0:
        out.put((rho_m, eta_m, w_eps, ids[0] == bytes(range(128))))
    else:
        assert ids[0] == bytes(range(128))
    dist.destroy_process_group()


def test_sharded_lse_merge_equals_unsharded_tick():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    rho_m, eta_m, w_eps, id_ok = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert id_ok
    g = Golden("diffdrive_pe0.05")
    K, T = 512, 30
    sp = orc.diffdrive_spec(K=K, T=T, param_exploration=0.05, cost_mode="sum", waypoint_mode="frozen")
    sp.temperature = 1.5
    x0, U = np.array([0.3, 0.2, 0.4]), np.zeros((T, 2))
    S, _, _ = co.costs(sp, g.path, U, 0, x0, None, seed=9, tick=2)
    eps = orc.philox_noise(9, 2, K, T, sp.sigma)
    o = orc.update_vec(sp, U, S, eps, 0)
    assert abs(rho_m - S.min()) < 1e-12
    assert np.max(np.abs(w_eps.reshape(T, 2) - o["w_eps"])) < 1e-10


# ---- the PRODUCT's host-side sharding logic under gloo (no GPU): how the drop-in class splits K over the ranks, what it
# hands the engine (K_local, K_global, k_offset -- the Philox counters and the Q6 explore split use the GLOBAL index) and the
# IPC-handle exchange of comm_init_from_torch.  The engine handle itself is replaced by a recorder: the CUDA side of the same
# calls is covered by tests/test_gpu_multi.py and by bench.py's parity_check on real GPUs.
class _RecorderEngine:
    made = []

    def __init__(self, **kw):
        self.kw = kw
        self.device = kw.get("device", 0)
        self.model = kw["model"]
        self.opened = None
        _RecorderEngine.made.append(self)

    def set_ref_path(self, p):
        self.path = np.asarray(p)

    def set_obstacles(self, o):
        pass

    def comm_p2p_export(self, world):
        return bytes([self.kw["k_offset"] % 251]) * 64          # a per-rank 64-byte "IPC handle"

    def comm_p2p_open(self, handles, rank, world):
        self.opened = (handles, rank, world)

    def close(self):
        pass


def _shard_worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "dnn-mppi-mpc_b200"), os.path.join(root, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mppi_b200._base as base
    base.MPPIEngine = _RecorderEngine                            # the class under test is the host logic around the handle
    from mppi_b200.mppi_differential_drive import MPPIAlgorithms
    g = Golden("diffdrive_pe0.05")
    K = 1 << 16
    ctrl = MPPIAlgorithms(delta_t=0.1, ref_path=g.path, max_speed=5.0, max_omega=3.14, num_samples_K=K, num_horizons_T=30,
                          param_exploration=0.05, param_lambda=1.0, param_alpha=0.2, sigma=np.array([[0.1, 0.0], [0.0, 0.01]]),
                          stage_cost_weight=np.array([5.0, 5.0, 10.0]), terminal_cost_weight=np.array([5.0, 5.0, 10.0]),
                          visualize_optimal_traj=False, visualze_sampled_trajs=False, cost_mode="sum", waypoint_mode="frozen",
                          temperature=2.0, seed=21, device=0, rank=rank, world=world)
    ctrl.comm_init_from_torch()
    e = ctrl.engine
    bad = None
    try:
        MPPIAlgorithms(delta_t=0.1, ref_path=g.path, max_speed=5.0, max_omega=3.14, num_samples_K=1001, num_horizons_T=30,
                       param_exploration=0.05, param_lambda=1.0, param_alpha=0.2, sigma=np.eye(2), stage_cost_weight=np.ones(3),
                       terminal_cost_weight=np.ones(3), visualize_optimal_traj=False, visualze_sampled_trajs=False,
                       rank=rank, world=world)
    except ValueError as ex:
        bad = str(ex)
    out.put((rank, dict(K=e.kw["K"], K_global=e.kw["K_global"], k_offset=e.kw["k_offset"], opened=e.opened, bad=bad,
                        K_attr=ctrl.K, T=e.kw["T"], temperature=e.kw["temperature"])))
    dist.barrier()
    dist.destroy_process_group()


def test_drop_in_class_shards_samples_by_global_index_and_exchanges_handles():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    K = 1 << 16
    for r in range(world):
        d = got[r]
        assert (d["K"], d["K_global"], d["k_offset"]) == (K // world, K, r * (K // world)), d
        assert d["K_attr"] == K and d["temperature"] == 2.0                  # the class still presents the GLOBAL sample count
        handles, rank, w = d["opened"]
        assert (rank, w) == (r, world) and len(handles) == 64 * world
        # every rank received every rank's handle, in rank order
        assert handles[:64] == bytes([0]) * 64 and handles[64:] == bytes([(K // world) % 251]) * 64
        assert d["bad"] and "divisible" in d["bad"]                          # K not divisible by the world size is refused
