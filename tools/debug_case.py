import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
name = sys.argv[1]
z = np.load(f"tests/golden/{name}.npz")
text = z["text"]; n = len(text)
obj = pkg.SuffixArray(text, idx_bytes=int(z["idx_bytes"])); obj.construct()
sa, lcp = obj.SA(), obj.LCP()
print(obj.stats())
bad = np.nonzero(lcp != z["lcp"])[0]
print("SA equal:", np.array_equal(sa, z["sa"]), "bad LCP count", len(bad))
for k in bad[:20]:
    a, b = int(sa[k - 1]), int(sa[k])
    print(f"k={k} SA[k-1]={a} SA[k]={b} got={lcp[k]} want={z['lcp'][k]} tail_a={bytes(text[a:a+40])} tail_b={bytes(text[b:b+40])}")
