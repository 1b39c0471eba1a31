"""Race hunting: python tools/debug_determinism.py [B] [distl 0/1] [critic|actor] — reports the
first launch of the prepared update whose outputs differ between two runs from the same state."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.parity import plan_divergence

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
distl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
which = sys.argv[3] if len(sys.argv) > 3 else "critic"
print(plan_divergence(B, distl, which, repeats=6) or "deterministic over 6 repeats")
