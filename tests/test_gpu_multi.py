"""bpc_sample_multi: several GPUs inside one process behind the C ABI (one host thread per device, ncclCommInitAll), driven from a
subprocess that never imports torch.  With one visible GPU the same entry point runs with a one-device communicator; the
two-device run needs `gpurun --gpus 2` (its log is kept under profiles/)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys
import numpy as np
import bpinference_b200 as bp
from bpinference_b200 import _ffi, synthetic as syn
assert "torch" not in sys.modules
L = _ffi.lib()
ndev = min(int(sys.argv[1]), L.bpc_device_count())
cfg = syn.make_config("cfg2", seed=1, chains=16, nobs=120)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"])
per = 16 // ndev
out = {"ndev": ndev, "nccl": L.bpc_nccl_version()}
# per-chain adaptation: chains are independent, so sharding must not change a single draw
one = eng.sampling(cfg["data"], chains=16, iter=200, control=dict(adapt_delta=0.85), seed=7, device=0)
multi = eng.sampling_multi(cfg["data"], list(range(ndev)), chains=per, iter=200, control=dict(adapt_delta=0.85), seed=7)
out["identical"] = bool((one.draws == multi.draws).all() and (one.n_leapfrog__ == multi.n_leapfrog__).all() and (one.lp__ == multi.lp__).all())
out["rhat_equal"] = bool(np.allclose(one.rhat, multi.rhat, rtol=1e-10))
out["total_chains"] = multi.info["total_chains"]
# pooled adaptation: the window statistics go through ncclAllReduce
pooled = eng.sampling_multi(cfg["data"], list(range(ndev)), chains=per, iter=400, control=dict(adapt_delta=0.85), seed=7, pooled_adaptation=True)
out["pooled_rhat_max"] = float(np.nanmax(pooled.rhat))
out["pooled_exchanges"] = pooled.info["pool_exchanges"]
ref = eng.sampling(cfg["data"], chains=16, iter=400, control=dict(adapt_delta=0.85), seed=7, pooled_adaptation=True, device=0)
z = (pooled.draws.reshape(-1, 6).mean(0) - ref.draws.reshape(-1, 6).mean(0)) / (ref.draws.reshape(-1, 6).std(0) / np.sqrt(np.minimum(ref.ess_bulk(), pooled.ess_bulk())))
out["pooled_vs_single_z"] = float(np.abs(z).max())
out["torch_loaded"] = "torch" in sys.modules
print("RESULT " + json.dumps(out))
"""


def _run(ndev):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", SCRIPT, str(ndev)], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_sample_multi_through_the_c_abi_without_torch():
    import torch
    want = 2 if torch.cuda.device_count() >= 2 else 1
    out = _run(want)
    assert out["ndev"] == want and out["nccl"] >= 20000 and not out["torch_loaded"]
    assert out["identical"] and out["rhat_equal"] and out["total_chains"] == 16
    assert out["pooled_rhat_max"] < 1.1 and out["pooled_exchanges"] >= 3
    assert out["pooled_vs_single_z"] < 6.0
    path = os.path.join(ROOT, "gpurun_out", "r2_sample_multi_%ddev.json" % want)
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        json.dump(out, open(path, "w"))
    except OSError:
        pass
