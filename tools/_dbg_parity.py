import sys, os, json
sys.path.insert(0, ".")
from tests import parity
for kw in (dict(seed=12, B=16384, distl=True, steps=1), dict(seed=11, B=8192, distl=False, steps=1)):
    r = parity.run_learner_parity(obs_dim=88, act_dim=16, device="cuda:0", check=False, **kw)
    pt = r.pop("per_tensor_grad (vs tf32 oracle, vs fp32 oracle, tf32 oracle vs fp32 oracle)")
    print(os.environ.get("PQLB_FWD_MODE"), kw, json.dumps(r))
    print("  v:", pt["v"]); print("  p:", pt["p"])
