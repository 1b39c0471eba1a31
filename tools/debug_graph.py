"""Debug helper: capture the first k prepared launches of a V-learner update into a CUDA graph,
replay, synchronise.  Usage: python tools/debug_graph.py K   (K = -1: driver loop over all k)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def probe(k, subset=None):
    import torch
    from tests import parity
    from tests.golden import inputs
    from pql_b200.algo import PQLVLearner
    from pql_b200.models import TanhMLPPolicy
    dev = torch.device("cuda:0")
    B, O, A = 512, 88, 16
    case = inputs.learner_case(1, B, O, A, False)
    cfg = parity.make_cfg(B, False)
    v = PQLVLearner(O, A, cfg)
    actor = TanhMLPPolicy(O, A).to(dev)
    v.update(actor, tuple(x.to(dev) for x in case["batch"]), None, 0)
    plan = v._plan
    torch.randint(B, (B,), device=dev, out=plan.idx)
    seq = [v._sample] + plan.calls + [plan.reduce_call, plan.adamw_call, plan.loss_call]
    names = [c.name for c in seq]
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    chosen = seq[:k] if subset is None else [seq[i] for i in subset]
    with torch.cuda.graph(g):
        for c in chosen:
            c()
    g.replay()
    torch.cuda.synchronize()
    print("OK", k, names[k - 1] if k else "-", flush=True)

if __name__ == "__main__":
    k = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("s") else -2
    if len(sys.argv) > 1 and sys.argv[1].startswith("s"):
        probe(len(sys.argv[1]), [int(x) for x in sys.argv[1][1:].split(",")])
    elif k == -2:
        for sub in ("s5", "s0,5", "s1,5", "s4,5", "s1,2,3,4,5", "s0,1,5", "s5,5", "s1,1,1,1,5", "s2,5", "s3,5", "s5,6,7"):
            r = subprocess.run([sys.executable, __file__, sub], capture_output=True, text=True)
            print(sub, r.returncode, (r.stdout.strip().splitlines() or ["FAIL"])[-1], flush=True)
    elif k >= 0:
        probe(k)
    else:
        for kk in range(1, 40):
            r = subprocess.run([sys.executable, __file__, str(kk)], capture_output=True, text=True)
            out = (r.stdout.strip().splitlines() or ["?"])[-1]
            print(kk, r.returncode, out, flush=True)
            if r.returncode != 0:
                print(r.stderr[-1500:])
                break
            if "adamw" in out and False:
                break
