"""Run every GPU test function in its own process (a device-side trap poisons the CUDA context of its process
only) with a per-process timeout, and write a summary to gpurun_out/isolated_summary.txt."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)


def main():
    files = sys.argv[1:] or ["tests/test_kernels_gpu.py"]
    out = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu", *files], cwd=ROOT,
                         capture_output=True, text=True).stdout
    funcs = []
    for line in out.splitlines():
        if "::" in line:
            f = line.split("[")[0]
            if f not in funcs:
                funcs.append(f)
    summary = []
    for f in funcs:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "--tb=short", "-p", "no:cacheprovider", f],
                               cwd=ROOT, capture_output=True, text=True, timeout=420)
            tail = "\n".join((r.stdout + r.stderr).splitlines()[-40:])
            status = "PASS" if r.returncode == 0 else f"FAIL(rc={r.returncode})"
        except subprocess.TimeoutExpired as e:
            status, tail = "TIMEOUT", str(e)[-2000:]
        summary.append(f"=== {f}: {status} ({time.time() - t0:.1f}s)\n{tail}\n")
        print(summary[-1], flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "isolated_summary.txt"), "w") as fh:
        fh.write("\n".join(summary))


if __name__ == "__main__":
    main()
