"""Ad-hoc fuzz of the exact chain kernels on the host emulation (tests/emu; build it first: python tests/emu/build_emu.py):
random N of 72-128 bits, depth and interval of 2^16..2^17.6 periods; the two-stage loop, the one-stage loop and the row
kernel must report the same divisors and counts.  python tools/emu_chain_fuzz.py [SEED] [SECONDS]"""
import importlib.util, os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qimcifa_b200 import abi as real_abi
spec = importlib.util.spec_from_file_location("qcf_emu_abi", real_abi.__file__)
abi = importlib.util.module_from_spec(spec); spec.loader.exec_module(abi)
abi.LIB_PATH = os.path.join(ROOT, 'tests', '_build', 'libqimcifa_emu.so')
from oracle import oracle_c as oc
import sympy
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
ctx = abi.Context(0)
t0=time.time(); ncase=0; twos=0
while time.time()-t0 < float(sys.argv[2] if len(sys.argv)>2 else 240):
    nbits = rnd.choice((72, 80, 96, 100, 112, 128))
    half = nbits // 2
    depth = rnd.choice((0, 0, 1, 2, 5))
    gb = max(34, half - depth)
    top = rnd.randrange(1 << (gb - 1), 1 << gb)
    periods = rnd.randrange(1 << 16, (1 << 17) + (1 << 16))
    level = rnd.choice((5, 5, 5, 6))
    period = 2310 if level == 5 else 30030
    if level == 6: periods = rnd.randrange(1<<16, (1<<16)+(1<<14))
    i_hi = oc.backward(top) - 1
    i_lo = max(2, oc.backward(max(17, top - periods * period)))
    v_lo, v_hi = oc.forward(i_lo), oc.forward(i_hi)
    if rnd.random() < 0.7:
        p = sympy.prevprime(rnd.randrange(v_lo + 1000, v_hi))
        cof_bits = max(2, nbits - p.bit_length())
        n = p * (rnd.randrange(1 << (cof_bits - 1), 1 << cof_bits) | 1)
    else:
        p = None
        n = rnd.randrange(1 << (nbits - 1), 1 << nbits) | 1
    if n.bit_length() > 128: continue
    out = {}
    for name, two, cm in (("two", "1", "16"), ("one", "0", "16"), ("rows", "0", "40")):
        abi.set_option("xchain_two_stage", two); abi.set_option("chain_min_log2", cm)
        r = ctx.sweep_indices(n, level, i_lo, i_hi, flags=abi.KERNEL_EXACT | abi.COUNT_ON_DEVICE)
        out[name] = (ctx.last_hits(), r.guesses_tested, r.guesses_counted, r.factor)
        if name == "two": l2 = r.kernel_launches
        if name == "rows": lr = r.kernel_launches
    assert out["two"] == out["one"] == out["rows"], (n, level, i_lo, i_hi, out)
    assert out["two"][1] == out["two"][2] == oc.intent_count(level, i_lo, i_hi)
    if p: assert p in out["two"][0]
    ncase += 1; twos += l2 > lr
    print(ncase, nbits, gb, level, periods, "chain launches" if l2 > lr else "rows only", len(out["two"][0]), "%.0fs" % (time.time()-t0), flush=True)
print("cases", ncase, "with chain launches", twos)
