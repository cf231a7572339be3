"""Run every conv/DCN kernel case on the GPU and report each one (development aid).

A kernel fault poisons the CUDA context, so cases run in child processes that
are restarted after the first faulting case.  Output: one line per case on
stdout and in gpurun_out/probe_conv.log.
"""
import json
import os
import subprocess
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(start, only):
    import torch
    from conv_cases import CASES, run_conv_case

    for i in range(start, len(CASES)):
        name, kw = CASES[i]
        if only and only not in name:
            continue
        try:
            err, scale, tol = run_conv_case(name, **kw)
            print(json.dumps(dict(i=i, name=name, ok=bool(err <= tol), err=err, scale=scale, tol=tol)), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(i=i, name=name, ok=False, exc=repr(e)[:400])), flush=True)
            traceback.print_exc()
            if "CUDA" in repr(e) or "cuda" in repr(e):
                return 3  # context is gone: let the parent restart after this case
    return 0


def main():
    from conv_cases import CASES

    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "probe_conv.log"), "w")
    start, results = 0, []
    while start < len(CASES):
        p = subprocess.run([sys.executable, __file__, "--child", str(start)] + (["--only", only] if only else []),
                           capture_output=True, text=True, timeout=900)
        last = start - 1
        for line in p.stdout.splitlines():
            if line.startswith("{"):
                r = json.loads(line)
                results.append(r)
                last = r["i"]
                print(line)
                log.write(line + "\n")
        if p.stderr.strip():
            log.write("---- stderr (child from %d)\n%s\n" % (start, p.stderr[-6000:]))
        if p.returncode == 0:
            break
        start = last + 1 if last >= start else start + 1
    bad = [r["name"] for r in results if not r["ok"]]
    summary = "%d/%d cases ok; failed: %s" % (len(results) - len(bad), len(results), bad)
    print(summary)
    log.write(summary + "\n")
    log.close()
    return 1 if bad else 0


if __name__ == "__main__":
    if "--child" in sys.argv:
        o = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""
        sys.exit(child(int(sys.argv[sys.argv.index("--child") + 1]), o))
    sys.exit(main())
