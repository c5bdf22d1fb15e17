"""Run GPU parity cases one subprocess at a time (a trapping kernel poisons only its own process).

  python tools/bringup.py [--cases a,b,c] [--timeout 120] [--out gpurun_out/bringup.jsonl]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_one(name):
    import conv_cases
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    r = conv_cases.CASES[name]()
    print("RESULT " + json.dumps(r))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="")
    ap.add_argument("--one", default="")
    ap.add_argument("--timeout", type=int, default=120)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bringup.jsonl"))
    a = ap.parse_args()
    if a.one:
        run_one(a.one)
        return
    import conv_cases
    names = [c for c in a.cases.split(",") if c] or list(conv_cases.CASES)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    n_ok = 0
    with open(a.out, "a") as f:
        for n in names:
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True,
                                   text=True, timeout=a.timeout)
                res = None
                for line in p.stdout.splitlines():
                    if line.startswith("RESULT "):
                        res = json.loads(line[7:])
                if res is None:
                    res = {"case": n, "ok": False, "rc": p.returncode, "stderr": p.stderr[-1500:]}
            except subprocess.TimeoutExpired:
                res = {"case": n, "ok": False, "timeout": True}
            res["name"] = n
            res["secs"] = round(time.time() - t0, 1)
            n_ok += bool(res.get("ok"))
            f.write(json.dumps(res) + "\n")
            f.flush()
            print(json.dumps(res))
    print("bringup: %d/%d ok" % (n_ok, len(names)))


if __name__ == "__main__":
    main()
