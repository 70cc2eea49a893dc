#!/usr/bin/env python
"""BASELINE config C1 on this host: 4096 files x 64 KiB of zpk-synth-v1 on disk, the UNMODIFIED reference CLI
(oracle/_ref/zpack_ref) and the same CLI sources linked against the drop-in library (oracle/_ref/dropin_zpack, GPU
behind lib/zpack.h) side by side: `c -m lz4`, `t`, `x -o` (and `c -m zstd` / `t`), best of 3, files in the page
cache.  Prints one JSON line per (cli, operation).  TEST / MEASUREMENT INFRASTRUCTURE: the only place the reference
binary is run next to the product."""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def main():
    from zpack_b200 import corpus
    n, size = 4096, 65536
    wd = tempfile.mkdtemp(prefix="zpb_c1_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    src = os.path.join(wd, "corpus")
    os.makedirs(src)
    for i in range(n):
        corpus.entry_bytes(i, size).tofile(os.path.join(src, f"f{i:05d}.bin"))
    total = n * size
    host = {"nproc": os.cpu_count(), "cpu": next((l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "?")}
    print(json.dumps({"host": host, "corpus": f"{n} files x {size} B zpk-synth-v1 in {wd}"}), flush=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "zpack_b200") + ":" + REFDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    for cli in ("zpack_ref", "dropin_zpack"):
        exe = os.path.join(REFDIR, cli)
        if not os.path.exists(exe):
            print(json.dumps({"cli": cli, "error": "not built"}))
            continue
        for method in ("lz4", "zstd"):
            arch = os.path.join(wd, f"{cli}_{method}.zpk")
            ops = [("c", [exe, "c", "-m", method, arch, src]), ("t", [exe, "t", arch]),
                   ("x", [exe, "x", "-o", os.path.join(wd, "out"), arch])]
            for name, cmd in ops:
                best, rc, tail = None, 0, ""
                for rep in range(3):
                    if name == "c" and os.path.exists(arch):
                        os.remove(arch)
                    shutil.rmtree(os.path.join(wd, "out"), ignore_errors=True)
                    t0 = time.perf_counter()
                    r = subprocess.run(cmd, cwd=wd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                    dt = time.perf_counter() - t0
                    rc, tail = r.returncode, r.stdout[-160:].replace("\n", " | ")
                    if rc != 0:
                        break
                    best = dt if best is None else min(best, dt)
                line = {"cli": cli, "method": method, "op": name, "rc": rc}
                if best is not None:
                    line.update({"best_of_3_s": round(best, 4), "uncompressed_GBps": round(total / best / 1e9, 3)})
                if name == "c" and os.path.exists(arch):
                    line["archive_bytes"] = os.path.getsize(arch)
                    line["ratio"] = round(total / os.path.getsize(arch), 3)
                if rc != 0:
                    line["tail"] = tail
                print(json.dumps(line), flush=True)
            if cli == "dropin_zpack" and os.path.exists(arch):     # the unmodified reference must accept what the drop-in wrote
                r = subprocess.run([os.path.join(REFDIR, "zpack_ref"), "t", arch], cwd=wd, env=env, capture_output=True, text=True)
                print(json.dumps({"cli": "zpack_ref", "op": f"t on the drop-in's {method} archive", "rc": r.returncode,
                                  "tail": r.stdout[-80:].replace("\n", " | ")}), flush=True)
    # fixed cost of a drop-in process (CUDA context + kernel setup): `t` on a one-file archive
    one = os.path.join(wd, "one")
    os.makedirs(one)
    corpus.entry_bytes(0, size).tofile(os.path.join(one, "f.bin"))
    for cli in ("zpack_ref", "dropin_zpack"):
        exe = os.path.join(REFDIR, cli)
        arch = os.path.join(wd, f"{cli}_one.zpk")
        subprocess.run([exe, "c", "-m", "lz4", arch, one], cwd=wd, env=env, capture_output=True)
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            r = subprocess.run([exe, "t", arch], cwd=wd, env=env, capture_output=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(json.dumps({"cli": cli, "op": "t on a one-file archive (process start-up)", "rc": r.returncode, "best_of_3_s": round(best, 4)}), flush=True)
    shutil.rmtree(wd, ignore_errors=True)


if __name__ == "__main__":
    main()
