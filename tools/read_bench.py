"""Host reader throughput: python tools/read_bench.py FILE.txt.gz  (EPI_INFLATE_THREADS / EPI_PARSE_THREADS / EPI_INFLATE_DEBUG=1)."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from epilogos_b200 import helpers
path = sys.argv[1]
ts = []
for rep in range(6):
    t0 = time.time(); loc, m = helpers.read_matrix(path, num_states=18); ts.append(time.time() - t0)
print("read_matrix min %.3f s median %.3f s -> %.0f k rows/s (min)" % (min(ts), sorted(ts)[len(ts)//2], m.shape[0] / 1e3 / min(ts)), m.shape)
