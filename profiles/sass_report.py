"""SASS evidence for the two Viterbi kernels: the loop bodies as disassembled from the built library, and an
opcode histogram per body (profiles/sass_r02/).  Also imported by tests/test_sass_budget.py, which fails when a
toolkit change silently loses the fused packed-min-with-predicates pattern or inflates the instruction budget.

    python profiles/sass_report.py          -> writes profiles/sass_r02/{viterbi_pair_acs_loop.sass, viterbi_warp_forward_loop.sass, summary.json}
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "viterbi.dll_b200", "libviterbi_b200.so")
INS = re.compile(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);")


def functions(lib=LIB):
    """{mangled name: [(address, text)]} for every kernel in the library."""
    out = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = INS.match(line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    return funcs


def loops(ins):
    """Backward branches as (target, branch address) pairs."""
    res = []
    for addr, text in ins:
        m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            res.append((int(m.group(1), 16), addr))
    return res


def opcode(text):
    text = re.sub(r"^@!?U?P\d+\s+", "@P ", text)
    parts = text.split()
    return "@P " + parts[1] if parts[0] == "@P" else parts[0]


def body_stats(ins, lo, hi):
    body = [(a, t) for a, t in ins if lo <= a <= hi]
    hist = collections.Counter(opcode(t) for _, t in body)
    return body, hist


def pair_kernel_loop(funcs, word_stores=True, punctured=False):
    """The 2-step ACS loop of viterbi_pair_kernel<kWordStores, kPunct>: the innermost loop that holds the butterflies."""
    tag = "ILb%dELb%dE" % (int(word_stores), int(punctured))
    name = next(n for n in funcs if "viterbi_pair_kernel" in n and tag in n)
    ins = funcs[name]
    best = None
    for lo, hi in loops(ins):
        body, hist = body_stats(ins, lo, hi)
        if hist.get("VIADDMNMX.U16x2", 0) >= 64 and (best is None or len(body) < len(best[0])):
            best = (body, hist)  # the innermost loop holding the butterflies
    return name, best[0], best[1]


def warp_kernel_loop(funcs):
    """The 10-step forward loop of viterbi_warp_kernel: the smallest loop with ten butterfly shuffles."""
    name = next(n for n in funcs if "viterbi_warp_kernel" in n and "ILi128E" in n)  # the latency shape (four warps per block)
    ins = funcs[name]
    best = None
    for lo, hi in loops(ins):
        body, hist = body_stats(ins, lo, hi)
        if hist.get("SHFL.BFLY", 0) >= 10 and (best is None or len(body) < len(best[0])):
            best = (body, hist)
    return name, best[0], best[1]


def summary(funcs=None):
    funcs = funcs or functions()
    res = {}
    name, body, hist = pair_kernel_loop(funcs)
    fused = sum(1 for _, t in body if re.match(r"VIMNMX\.U16x2 R\d+, P\d, P\d,", t))
    res["viterbi_pair_kernel"] = {
        "function": name, "loop_instructions": len(body), "trellis_steps_per_iteration": 2,
        "instructions_per_step_64_frames": len(body) / 2.0, "vimnmx_u16x2_with_two_predicate_outputs": fused,
        "histogram": dict(hist.most_common())}
    pname, pbody, phist = pair_kernel_loop(funcs, punctured=True)
    res["viterbi_pair_kernel_punctured"] = {"function": pname, "loop_instructions": len(pbody),
                                            "extra_instructions_vs_plain": len(pbody) - len(body),
                                            "histogram": dict(phist.most_common())}
    name, wbody, whist = warp_kernel_loop(funcs)
    res["viterbi_warp_kernel"] = {"function": name, "loop_instructions": len(wbody), "trellis_steps_per_iteration": 10,
                                  "instructions_per_step": len(wbody) / 10.0, "histogram": dict(whist.most_common())}
    return res, body, wbody


def main():
    res, body, wbody = summary()
    out = os.path.join(ROOT, "profiles", "sass_r02")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "viterbi_pair_acs_loop.sass"), "w") as f:
        f.write("// viterbi_pair_kernel<true>: the ACS loop body (2 trellis steps of 64 frames per warp), cuobjdump -sass of the built library\n")
        f.writelines("/*%04x*/  %s ;\n" % (a, t) for a, t in body)
    with open(os.path.join(out, "viterbi_warp_forward_loop.sass"), "w") as f:
        f.write("// viterbi_warp_kernel: the forward loop body (10 trellis steps of one frame), cuobjdump -sass of the built library\n")
        f.writelines("/*%04x*/  %s ;\n" % (a, t) for a, t in wbody)
    with open(os.path.join(out, "summary.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "histogram"} for k, v in res.items()}, indent=1))


if __name__ == "__main__":
    sys.exit(main())
