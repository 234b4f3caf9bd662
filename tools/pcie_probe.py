"""PCIe floor for the host-buffer (`e2e`) path: pinned H2D, D2H and both at once, in the chunk sizes run_chunked uses, next to
the e2e rate of the library at a few ROGTK_CHUNK_BYTES settings.  Run on the GPU box:  python tools/pcie_probe.py [n_umis]"""
import json
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def copy_rates(total=8 << 30, chunk=256 << 20):
    h_in = torch.empty(total, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(total, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(chunk * 2, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(chunk * 2, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    n = total // chunk
    res = {}

    def run(do_in, do_out):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            if do_in:
                with torch.cuda.stream(s1):
                    d_in[(i & 1) * chunk:(i & 1) * chunk + chunk].copy_(h_in[i * chunk:(i + 1) * chunk], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    h_out[i * chunk:(i + 1) * chunk].copy_(d_out[(i & 1) * chunk:(i & 1) * chunk + chunk], non_blocking=True)
        torch.cuda.synchronize()
        return total / (time.perf_counter() - t0) / 1e9

    for _ in range(2):
        res["h2d_GBps"] = run(True, False)
        res["d2h_GBps"] = run(False, True)
        res["both_each_GBps"] = run(True, True)
    return res


def main():
    n_umis = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    out = {"copy": copy_rates()}
    print(json.dumps(out), flush=True)
    for cb in (64 << 20, 128 << 20, 256 << 20, 512 << 20):
        env = dict(os.environ, ROGTK_CHUNK_BYTES=str(cb))
        p = subprocess.run([sys.executable, "bench.py", "--umis", str(n_umis), "--no-cpu", "--steps", "2", "--warmup", "3",
                            "--e2e-steps", "3"], env=env, capture_output=True, text=True)
        try:
            e = json.loads(p.stdout.strip().splitlines()[-1])["e2e"]
            print(json.dumps({"chunk_bytes": cb, "e2e_ms": e["ms_per_step"], "e2e_obs_per_s": e["value"],
                              "h2d_GBps": e["h2d_bytes_per_step"] / e["ms_per_step"] / 1e6,
                              "d2h_GBps": e["d2h_bytes_per_step"] / e["ms_per_step"] / 1e6}), flush=True)
        except Exception:
            print(json.dumps({"chunk_bytes": cb, "error": p.stderr[-400:]}), flush=True)


if __name__ == "__main__":
    main()
