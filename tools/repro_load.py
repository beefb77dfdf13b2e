import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import khmer_b200 as kh
d = "tests/golden/data/"
print("A", flush=True)
ng = kh.Nodegraph.load(d + "goodversion-k12.ht")
print("B", ng.ksize(), flush=True)
c = kh.Countgraph(1, 1, 1, primes=[1])
print("C created", flush=True)
del c
print("C deleted", flush=True)
try:
    kh.Countgraph.load(d + "goodversion-k12.ht")
except Exception as e:
    print("D", type(e), e, flush=True)
print("E", flush=True)
