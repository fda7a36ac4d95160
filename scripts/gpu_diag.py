"""Runs every GPU test in its OWN subprocess with a timeout (a trapped kernel poisons the CUDA
context of its process; a subprocess per test keeps the rest of the run meaningful) and writes a
summary to gpurun_out/diag_summary.txt plus the full logs to gpurun_out/diag/.

    python scripts/gpu_diag.py [pytest -k expression] [--timeout 180]
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    args = sys.argv[1:]
    timeout = 180
    if "--timeout" in args:
        i = args.index("--timeout")
        timeout = int(args[i + 1])
        del args[i:i + 2]
    files = [a for a in args if a.endswith(".py")] or ["tests/test_gpu_kernels.py", "tests/test_gpu_dit.py"]
    kexpr = [a for a in args if not a.endswith(".py")]
    out_dir = os.path.join(ROOT, "gpurun_out", "diag")
    os.makedirs(out_dir, exist_ok=True)
    collect = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu"] + files +
                             (["-k", kexpr[0]] if kexpr else []), cwd=ROOT, capture_output=True, text=True)
    ids = [l.strip() for l in collect.stdout.splitlines() if "::" in l]
    summary = []
    t_all = time.time()
    for tid in ids:
        safe = tid.replace("/", "_").replace("::", "__").replace("[", "_").replace("]", "")[:150]
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "--no-header", "-p", "no:cacheprovider",
                                tid], cwd=ROOT, capture_output=True, text=True, timeout=timeout)
            status = "PASS" if r.returncode == 0 else f"FAIL({r.returncode})"
            log = r.stdout + "\n" + r.stderr
        except subprocess.TimeoutExpired as ex:
            status = "TIMEOUT"
            log = (ex.stdout or b"").decode("utf-8", "replace") + "\n" + (ex.stderr or b"").decode("utf-8", "replace")
        dt = time.time() - t0
        with open(os.path.join(out_dir, safe + ".log"), "w") as f:
            f.write(log)
        tail = ""
        if status != "PASS":
            lines = [l for l in log.splitlines() if l.strip()]
            keep = [l for l in lines if ("Error" in l or "assert" in l or "err" in l.lower())][-6:]
            tail = " | " + " ; ".join(k.strip()[:160] for k in keep)
        line = f"{status:12s} {dt:6.1f}s {tid}{tail}"
        print(line, flush=True)
        summary.append(line)
    with open(os.path.join(ROOT, "gpurun_out", "diag_summary.txt"), "w") as f:
        f.write("\n".join(summary) + f"\n# total {time.time() - t_all:.0f}s, {len(ids)} tests\n")
    n_bad = sum(1 for s in summary if not s.startswith("PASS"))
    print(f"diag: {len(ids) - n_bad}/{len(ids)} passed")
    return 0


if __name__ == "__main__":
    sys.exit(main())
