"""Wall time of the arbplf-deriv executable on the cfg2 document (stdin -> stdout), with and without the warm-up thread."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
doc, N = bench.model_document(64)
codes = np.random.default_rng(0).integers(0, 4, (S, N)).astype(np.uint8)
for a, b in doc["model_and_data"]["edges"]:
    codes[:, a] = 4
text = bench.json_document_bytes(doc, codes)
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "phyly_b200", "bin", "arbplf-deriv")
for rep in range(3):
    for warm in (1, 0):
        env = dict(os.environ)
        if not warm:
            env["ARBPLF_NO_WARMUP"] = "1"
        t0 = time.perf_counter()
        pr = subprocess.run([exe], input=text, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        print("warmup=%d wall %.3f s rc %d" % (warm, time.perf_counter() - t0, pr.returncode))
