"""world_size-2 gloo test (CPU): the row-block sharding + the three statistics all-reduces of the N>1 path.

The CUDA kernels cannot run here, so each rank computes ITS SHARD's statistics with numpy from the packed layout
(the same quantities the kernels put in red1 / red3), all-reduces them with torch.distributed (gloo) exactly like
`CaviEngine._allreduce`, and rank 0 compares with the oracle's global values."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.golden_util import Golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vimure_b200 as vm
    from vimure_b200 import _packing
    from vimure_b200.model import shard_rows

    g = Golden(name)
    spec = g.R_spec
    mask = vm.masks.EgoMask(g.L, g.N, g.M, rep=spec["rep"], diag=spec["diag"])
    row0, nloc = shard_rows(g.N, world, rank)
    P = _packing.pack(g.X_subs, g.X_vals, g.L, g.N, g.M, g.K, mask, "cpu", row0=row0, nloc=nloc, tile_h=16)
    t = {k: v.numpy() for k, v in P.t.items()}
    st = g.init_state()
    # prior of this rank's special ties (rho = pr_rho at iteration 0)
    flat = t["u_gflat"]
    pr_u = np.zeros((P.U, g.K))
    pr_u[:, 0] = 1.0
    tf = (st["pr_ties"][:, 0] * g.N + st["pr_ties"][:, 1]) * g.N + st["pr_ties"][:, 2]
    o = np.argsort(tf)
    pos = np.minimum(np.searchsorted(tf[o], flat), len(tf) - 1)
    hit = tf[o][pos] == flat
    pr_u[hit] = st["pr_vals"][o][pos[hit]]
    # --- red3-like: this shard's share of A[l,m,k] = sum of rho_k over the ties reported by (l,m)
    L, N, M, K = g.L, g.N, g.M, g.K
    rep = spec["rep"].astype(bool)
    A = np.zeros((L, M, K))
    rows = np.arange(row0, row0 + nloc)
    for l in range(L):
        for m in np.nonzero(rep[l])[0]:
            n_loc = (N if row0 <= m < row0 + nloc else 0) + nloc - ((1 if spec["diag"] else 2) if row0 <= m < row0 + nloc else 0)
            A[l, m, 0] = n_loc
    ul = t["u_lrow"] // nloc
    ui = t["u_lrow"] % nloc + row0
    uj = t["u_col"]
    delta = pr_u.copy()
    delta[:, 0] -= 1.0
    for u in range(P.U):
        l, i, j = ul[u], ui[u], uj[u]
        if i == j:
            if spec["diag"] and i < M and rep[l, i]:
                A[l, i] += delta[u]
            continue
        if i < M and rep[l, i]:
            A[l, i] += delta[u]
        if j < M and rep[l, j]:
            A[l, j] += delta[u]
    # --- red1-like: this shard's share of sum over X entries of (l,m) of x
    g1 = np.zeros((L, M))
    el = t["u_lrow"][t["e_u"]] // nloc
    np.add.at(g1, (el, t["e_m"]), t["e_x"].astype(np.float64))
    A_t, g1_t = torch.from_numpy(A), torch.from_numpy(g1)
    dist.all_reduce(A_t)
    dist.all_reduce(g1_t)
    cnt = torch.tensor([P.I, P.IT, P.U], dtype=torch.int64)
    dist.all_reduce(cnt)
    if rank == 0:
        np.savez(out, A=A_t.numpy(), g1=g1_t.numpy(), cnt=cnt.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["f1_over", "gm_l2_k3"])
def test_two_rank_statistics_match_global(name, tmp_path):
    from oracle.cavi_numpy import OracleCAVI

    out = str(tmp_path / "res.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, name, out), nprocs=2, join=True)
    z = np.load(out)
    g = Golden(name)
    o = OracleCAVI(g.L, g.N, g.M, g.K, g.X_subs, g.X_vals, g.R_spec, mutuality=g.mutuality, **g.priors())
    st = g.init_state()
    o.set_state(st["gamma_shp"], st["gamma_rte"], st["phi_shp"], st["phi_rte"], st["nu_shp"],
                o.default_pr_rho(st["pr_ties"], st["pr_vals"]))
    A_ref = np.stack([o._reporter_sums(o.rho[..., k]) for k in range(g.K)], axis=-1)
    np.testing.assert_allclose(z["A"], A_ref, rtol=1e-12, atol=1e-12)
    g1_ref = np.zeros((g.L, g.M))
    np.add.at(g1_ref, (g.X_subs[0], g.X_subs[3]), g.X_vals.astype(float))
    np.testing.assert_allclose(z["g1"], g1_ref, rtol=0, atol=0)
    assert int(z["cnt"][0]) == len(g.X_vals)


def test_shard_rows_partition():
    from vimure_b200.model import shard_rows

    for N in (1, 7, 100, 20000, 64001):
        for W in (1, 2, 3, 8):
            cover = []
            for r in range(W):
                r0, n = shard_rows(N, W, r)
                cover += list(range(r0, r0 + n)) if N < 1000 else [(r0, n)]
            if N < 1000:
                assert cover == list(range(N))
            else:
                assert sum(n for _, n in cover) == N and all(cover[i][0] + cover[i][1] == cover[i + 1][0] for i in range(W - 1))


def _exchange_worker(rank, world, port, name, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vimure_b200.model import _exchange_reciprocals, shard_rows

    g = Golden(name)
    s, v = g.X_subs, g.X_vals
    row0, nloc = shard_rows(g.N, world, rank)
    own = (s[1] >= row0) & (s[1] < row0 + nloc)
    subs, vals = _exchange_reciprocals(tuple(s[:, own]), v[own], g.N, world, rank, torch.device("cpu"))
    got = np.stack([t.numpy() for t in (*subs, vals)], axis=1)
    np.save(out + ".%d.npy" % rank, got)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ingestion_exchange_of_reciprocal_entries(world, tmp_path):
    """`fit(presharded="rows")`: from the entries of its own rows every rank must end up with exactly the entries a
    row-block shard needs -- its own rows plus every entry whose column node it owns -- each once."""
    from vimure_b200.model import shard_rows

    name = "gm_l2_k3"
    out = str(tmp_path / "ex")
    mp.spawn(_exchange_worker, args=(world, _free_port(), name, out), nprocs=world, join=True)
    g = Golden(name)
    s, v = g.X_subs, g.X_vals
    full = np.concatenate([s.T, v[:, None]], axis=1).astype(np.int64)
    for rank in range(world):
        row0, nloc = shard_rows(g.N, world, rank)
        own = (s[1] >= row0) & (s[1] < row0 + nloc)
        tr = (s[2] >= row0) & (s[2] < row0 + nloc)
        want = np.concatenate([full[own], full[tr & ~own]], axis=0)
        got = np.load(out + ".%d.npy" % rank).astype(np.int64)
        key = lambda a: a[np.lexsort(a.T[::-1])]  # noqa: E731
        np.testing.assert_array_equal(key(got), key(want))
