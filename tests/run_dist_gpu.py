"""Multi-GPU check, to be launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_dist_gpu.py

Every rank fits the same golden scenarios through the public API with row-block sharding + NCCL all-reduces and
compares with the reference's golden trajectory (which a single-rank fit reproduces too)."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tests.golden_util import Golden  # noqa: E402
from tests.test_gpu_parity import build_inputs  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import vimure_b200 as vm

    ok = True
    for name in ("f1_over", "gm_l2_k3", "karnataka_vil1", "dense_reporting", "custom_mask"):
        g = Golden(name)
        X, R = build_inputs(g)
        mk = dict(g.model_kwargs)
        mk["convergence_tol"] = 0.0
        model = vm.VimureModel(**mk)
        fk = dict(g.fit_kwargs)
        n_it = min(g.n_iter, 20)
        fk["max_iter"] = n_it
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model.fit(X, R=R, init_state=g.init_state(), **fk)
        z = g.z
        it = n_it - 1
        try:
            np.testing.assert_allclose(model.gamma_shp, z["it_gamma_shp"][it], rtol=1e-5)
            np.testing.assert_allclose(model.gamma_rte, z["it_gamma_rte"][it], rtol=1e-5)
            np.testing.assert_allclose(model.phi_shp, z["it_phi_shp"][it], rtol=1e-5)
            np.testing.assert_allclose(model.phi_rte, z["it_phi_rte"][it], rtol=1e-5)
            np.testing.assert_allclose(model.maxL, z["it_elbo"][it], rtol=1e-6)
            rho = model.rho  # collective gather
            if "rho_final" in z.files and n_it == g.n_iter:
                np.testing.assert_allclose(rho, z["rho_final"], rtol=2e-5, atol=1e-30)
            print(f"[rank {rank}/{world}] {name}: OK  elbo={model.maxL:.6f}")
        except AssertionError as e:
            ok = False
            print(f"[rank {rank}/{world}] {name}: FAIL {str(e)[:300]}")
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if int(t.item()) == 1 else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
