"""Quick on-GPU probe: integer-pipe peaks + exact-kernel throughput per config (not a bench)."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qimcifa_b200 import abi

out = {}
with abi.Context(0) as ctx:
    out["device"] = ctx.device_info()
    for kind, name in ((0, "imad"), (1, "imad_wide"), (2, "iadd")):
        best = max(ctx.microbench(kind, 1 << 13)[0] for _ in range(3))
        out[name + "_ops_per_s"] = best
    flags = int(os.environ.get("QCF_KIND", "0"))
    for nbits, n in ((64, (2**32 - 5) * (2**32 - 17)), (96, (2**48 - 59) * (2**48 - 65)),
                     (112, (2**56 - 5) * (2**56 - 27)), (128, (2**64 - 59) * (2**64 - 83))):
        _, _, r, bc = abi.node_split(n, 1)
        for level in (5, 6, 7, 8, 9):
            nb = min(bc, 16)
            sec, g = ctx.time_level(n, level, nb, flags)
            sec, g = ctx.time_level(n, level, nb, flags)
            out[f"N{nbits}_L{level}"] = {"batches": nb, "guesses": g, "sec": sec, "gps": g / sec}
            print(nbits, level, nb, g, sec, "%.3e" % (g / sec), flush=True)
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
