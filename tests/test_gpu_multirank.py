"""N>1 parity ON GPUS (the reference asserts rank-count invariance: tests/test_python_repro_allegro.py:44-47, 302-327):
2 (and 4) ranks, one per GPU, brick domains, ghost positions / forces exchanged by the product halo code (alg_comm_*,
grouped ncclSend/ncclRecv) -- owned forces and per-atom energies equal the single-GPU result of the same box, the
summed energy / virial equal the single-box values, and every rank is bit-reproducible run to run.
Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`); skipped on a 1-GPU box."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,lmax,nlayers", [(2, 1, 2), (2, 2, 2), (4, 1, 2)])
def test_multi_gpu_equals_single_gpu(world, lmax, nlayers, ensure_built):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(29700 + world + lmax), os.path.join(ROOT, "tests", "mr_worker.py"), tmp, "8", str(lmax), str(nlayers)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        one = np.load(os.path.join(tmp, "single.npz"))
        fN, eN, seen = np.zeros_like(one["f"]), np.zeros_like(one["e"]), np.zeros(len(one["e"]), dtype=int)
        eng, halo = 0.0, 0.0
        for k in range(world):
            z = np.load(os.path.join(tmp, "rank%d.npz" % k))
            fN[z["tag"] - 1] = z["f"]; eN[z["tag"] - 1] = z["e"]; seen[z["tag"] - 1] += 1
            eng += float(z["eng"]); halo += float(z["halo"][0])
            np.testing.assert_allclose(z["tot"], one["tot"], rtol=1e-5, atol=1e-5)      # allreduce'd energy + virial == single box
    assert np.all(seen == 1) and halo > 0                      # every atom owned once; positions really crossed NVLink
    assert np.abs(fN - one["f"]).max() < 2e-5                  # fp32 model, different edge order per rank
    np.testing.assert_allclose(eN, one["e"], rtol=1e-5, atol=1e-5)
    assert abs(eng - float(one["eng"])) < 1e-6 * max(1.0, abs(float(one["eng"])))
