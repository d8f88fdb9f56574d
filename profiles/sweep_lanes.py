import os, sys, json, subprocess
sys.path.insert(0, '/root/repo')
res = []
for cfg in ("30 2 4096 20261020 -1", "100 3 256 20261021 -1"):
    for pers in ("1", "2", "0"):
        for warps in ("8", "16"):
            for chunk in ("64", "128", "256", "512"):
                env = dict(os.environ, SNAQ_LANES_PERSISTENT=pers, SNAQ_LANES_WARPS=warps, SNAQ_LANES_CHUNK=chunk)
                out = subprocess.run([sys.executable, "profiles/run_runs.py"] + cfg.split(), env=env, capture_output=True, text=True).stdout.strip().splitlines()
                try:
                    r = json.loads(out[-1]); print(cfg.split()[0], "pers", pers, "warps", warps, "chunk", chunk, "ms %.4f" % r["ms"], flush=True)
                except Exception as e:
                    print(cfg, pers, warps, chunk, "failed", out[-1:] , flush=True)
