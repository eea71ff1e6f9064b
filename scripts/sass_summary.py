"""Codegen evidence for profiles/: registers, spills, shared memory and SASS instruction mix of every hot kernel of
libldpc_b200.so, from `cuobjdump -res-usage` and `cuobjdump -sass` (run here: no GPU needed).

  python scripts/sass_summary.py [lib] > profiles/r06_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ldpc_decoder_b200", "libldpc_b200.so")
MNEMONICS = ["LDG.E.EF.128", "STG.E.EF.128", "LDG.E.128.CONSTANT", "LDG.E.CONSTANT", "LDS.U16", "LDS.128", "STS.128",
             "HMNMX2", "HADD2", "HFMA2", "PRMT", "LOP3.LUT", "MUFU.EX2", "MUFU.LG2", "FFMA", "FADD", "FMUL", "SHFL",
             "UBLKCP", "SYNCS", "ATOMG", "REDG", "BAR.SYNC", "BRX"]
HOT = [r"cn_update_kernel<.*TraitsF16, 6, unsigned int>", r"cn_update_kernel<.*TraitsF32, 6, unsigned int>",
       r"vn_update_kernel<.*TraitsF16, false, 6, unsigned int>", r"vn_update_kernel<.*TraitsF16, true, 6, unsigned int>",
       r"vn_update_kernel<.*TraitsF32, false, 6, unsigned int>", r"vn_update_kernel<.*TraitsF32, false, 4, unsigned int>",
       r"cn_update_kernel<.*TraitsF16, 6, unsigned long>", r"vn_update_kernel<.*TraitsF16, false, 6, unsigned long>",
       r"cn_update_refill_kernel<.*TraitsF16, 6, unsigned int>", r"cn_update_split_kernel<.*TraitsF32, 4>",
       r"cn_update_bulk_kernel<.*TraitsF16, 6>", r"cn_update_bulk_kernel<.*TraitsF32, 6>",
       r"refill_kernel<.*TraitsF16, 0, false>", r"refill_kernel<.*TraitsF16, 0, true>", r"refill_kernel<.*TraitsF32, 2, true>",
       r"refill_syndrome_kernel", r"syndrome_check_kernel", r"extract_kernel", r"sched_kernel", r"move_slots_kernel<float>"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"ldpc_b200::dev::", "", re.sub(r"\(.*", "", o.replace("(bool)1", "true").replace("(bool)0", "false")
                                                      .replace("(int)", ""))) for o in out]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", ln)
        if m and cur:
            usage[cur] = tuple(int(x) for x in m.groups())
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts, order = {}, []
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + ".") or (mn.endswith(".128") and op.startswith(mn)):
                    counts[cur][mn] += 1
    names = demangle(order)
    print(f"# SASS / resource summary of {os.path.relpath(LIB, ROOT)}  (cuobjdump -res-usage / -sass; architectures in the file: {', '.join(arch)})")
    print("# REG = registers per thread, STACK = bytes of local memory (0 = no spill frame), SHARED = static shared memory;")
    print("# instruction counts are STATIC counts over the whole kernel (all degree-specialised paths), not per-edge dynamic counts.\n")
    for pat in HOT:
        for mangled, name in zip(order, names):
            if re.search(pat, name):
                reg, stack, shared = usage.get(mangled, (-1, -1, -1))
                c = counts[mangled]
                mix = ", ".join(f"{mn} x{c[mn]}" for mn in MNEMONICS if c[mn])
                print(f"{name}\n    REG {reg}  STACK {stack}  SHARED {shared}  SASS instructions {c['_total']}\n    {mix}\n")
    # all kernels, one line each
    print("# every kernel in the library: REG / STACK / SASS instructions")
    for mangled, name in sorted(zip(order, names), key=lambda x: x[1]):
        reg, stack, shared = usage.get(mangled, (-1, -1, -1))
        print(f"{reg:4d} {stack:4d} {counts[mangled]['_total']:6d}  {name}")
    tens = sum(c[k] for c in counts.values() for k in c if k.startswith(("UTC", "HMMA", "IMMA", "LDTM")))
    print(f"\n# tensor-core instructions (UTC*MMA / HMMA / LDTM) in the library: {tens} -- sparse message passing is not a contraction")


if __name__ == "__main__":
    main()
