"""The sharded path (row blocks over ranks + all-reduced statistics) driven through `VimureModel.fit` by TWO ranks.
With two GPUs the ranks use NCCL on their own devices; with one GPU they share it and reduce with gloo (the kernels never
wait on one another, so two processes on one device are safe).  Covers what the advisor flagged: ranks must agree on the
RNG with seed=None, on the undirected-symmetry error and on the stop decision."""
import os
import socket
import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ngpu, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    dev = rank if ngpu >= world else 0
    torch.cuda.set_device(dev)
    if ngpu >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import vimure_b200 as vm
    from tests.golden_util import Golden
    from tests.test_gpu_parity import build_inputs

    res = {}
    tdev = "cuda" if ngpu >= world else "cpu"  # NCCL gathers device tensors, gloo host tensors
    try:
        # 1. golden trajectories through the sharded fit (injected state), incl. a general mask and N >= 512
        for name in ("f1_over", "gm_l2_k3", "sbm_n520", "custom_mask"):
            g = Golden(name)
            X, R = build_inputs(g)
            mk = dict(g.model_kwargs)
            mk["convergence_tol"] = 0.0
            model = vm.VimureModel(**mk)
            fk = dict(g.fit_kwargs)
            n_it = min(g.n_iter, 12)
            fk["max_iter"] = n_it
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                model.fit(X, R=R, init_state=g.init_state(), **fk)
            z, it = g.z, n_it - 1
            err = 0.0
            for k in ("gamma_shp", "gamma_rte", "phi_shp", "phi_rte"):
                err = max(err, float(np.max(np.abs(getattr(model, k) - z["it_" + k][it]) / np.abs(z["it_" + k][it]))))
            res[name] = (err, abs(model.maxL - float(z["it_elbo"][it])) / abs(float(z["it_elbo"][it])), model._world)
            if name == "f1_over":  # consumers on a sharded fit: every rank gets the whole (L, N, N) result
                Y = model.get_inferred_model("rho_max")
                res["rho_max_equal"] = bool(np.array_equal(Y, np.argmax(model.rho_f, axis=-1)))
        # 2. seed=None: the ranks must draw the same state, run the same number of iterations and agree on everything
        g = Golden("f1_over")
        X, R = build_inputs(g)
        model = vm.VimureModel(mutuality=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model.fit(X, R=R, K=2, seed=None, max_iter=40, num_realisations=2)
        t = torch.tensor([model.maxL, float(model.n_iter_), float(model.gamma_shp.sum()), float(model.nu_shp)],
                         dtype=torch.float64, device=tdev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        res["seed_none_agree"] = bool(all(torch.equal(p, parts[0]) for p in parts))
        # 3. undirected, init="fast": symmetric prior across ranks, same state on every rank
        g = Golden("undirected")
        X, R = build_inputs(g)
        model = vm.VimureModel(undirected=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model.fit(X, R=R, K=2, seed=7, max_iter=11, init="fast")
        t = torch.tensor([model.maxL, float(model.gamma_shp.sum())], dtype=torch.float64, device=tdev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        res["undirected_fast_agree"] = bool(all(torch.equal(p, parts[0]) for p in parts)) and bool(np.isfinite(model.maxL))
        # 4. an asymmetric X with undirected=True raises on EVERY rank (no rank is left waiting in a collective)
        s = g.X_subs.copy()
        keep = ~((s[1] == s[1][0]) & (s[2] == s[2][0]))  # drop one direction of one pair
        Xa = vm.sptensor.sptensor(tuple(s[:, keep]), g.X_vals[keep], shape=(g.L, g.N, g.N, g.M))
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                vm.VimureModel(undirected=True).fit(Xa, R=R, K=2, seed=7, max_iter=3)
            res["asym_raises"] = False
        except ValueError:
            res["asym_raises"] = True
        # 5. sharded ingestion: the whole list on every rank, own rows + reciprocals (presharded=True) and own rows only
        #    (presharded="rows": the reciprocals come from their owners in one all-to-all) give the same fit
        from vimure_b200.model import shard_rows

        g = Golden("sbm_n520")
        X, R = build_inputs(g)
        s, v = np.stack(X.subs), np.asarray(X.vals)
        row0, nloc = shard_rows(g.N, world, rank)
        own = (s[1] >= row0) & (s[1] < row0 + nloc)
        tr = (s[2] >= row0) & (s[2] < row0 + nloc)
        fits = {}
        for tag, sel, ps in (("all", np.ones(len(v), bool), False), ("own+reciprocals", own | tr, True), ("rows", own, "rows")):
            Xs = vm.sptensor.sptensor(tuple(s[:, sel]), v[sel], shape=X.shape)
            m = vm.VimureModel(mutuality=True, convergence_tol=0.0)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                m.fit(Xs, R=R, K=g.K, seed=3, max_iter=11, init="fast", presharded=ps)
            fits[tag] = (m.maxL, m.gamma_shp.copy(), m.phi_rte.copy(), float(m.nu_shp))
        res["ingestion_modes_agree"] = all(
            fits[t][0] == fits["all"][0] and np.array_equal(fits[t][1], fits["all"][1]) and
            np.array_equal(fits[t][2], fits["all"][2]) and fits[t][3] == fits["all"][3] for t in ("own+reciprocals", "rows"))
        res["ok"] = True
    except Exception as e:  # noqa: BLE001
        import traceback

        res["ok"] = False
        res["error"] = traceback.format_exc()[-1500:]
    out.put((rank, res))
    dist.destroy_process_group()


def test_two_rank_sharded_fit_matches_the_reference():
    import torch
    import torch.multiprocessing as mp

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    ngpu, world = torch.cuda.device_count(), 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, ngpu, out)) for r in range(world)]
    for p in ps:
        p.start()
    results = dict(out.get(timeout=900) for _ in ps)
    for p in ps:
        p.join(timeout=120)
    for rank in range(world):
        res = results[rank]
        assert res["ok"], res.get("error")
        for name in ("f1_over", "gm_l2_k3", "sbm_n520", "custom_mask"):
            err, e_elbo, w = res[name]
            assert w == world
            assert err <= 1e-5 and e_elbo <= 1e-6, (rank, name, err, e_elbo)
        assert res["rho_max_equal"] and res["seed_none_agree"] and res["undirected_fast_agree"] and res["asym_raises"], res
        assert res["ingestion_modes_agree"], res
