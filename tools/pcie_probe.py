"""Developer probe: pinned-host PCIe bandwidth of the box, each direction alone and both together.
`--all` runs one process per visible GPU at the same time (the platform's aggregate, what an N-GPU e2e can reach)."""
import os, subprocess, sys, time
import torch


def one(tag=""):
    n = 2 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=4):
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        return reps * n / dt / 1e9
    run(True, True, 1)
    start = float(os.environ.get("PROBE_START", "0"))
    while time.time() < start:
        time.sleep(0.01)
    a = run(True, False, 6); b = run(False, True, 6); c = run(True, True, 6)
    print(f"{tag}H2D alone {a:.1f} GB/s | D2H alone {b:.1f} GB/s | both: each direction {c:.1f} GB/s", flush=True)


if __name__ == "__main__":
    if "--all" in sys.argv:
        n = torch.cuda.device_count()
        env0 = dict(os.environ, PROBE_START=str(time.time() + 25))
        ps = [subprocess.Popen([sys.executable, __file__], env=dict(env0, CUDA_VISIBLE_DEVICES=str(i), PROBE_TAG=f"gpu{i}: ")) for i in range(n)]
        for p in ps:
            p.wait()
    else:
        one(os.environ.get("PROBE_TAG", ""))
