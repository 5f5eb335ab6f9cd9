"""Mnemonic counts and a short main-loop excerpt of the hot kernels in libcalclens_b200.so (cuobjdump -sass):
   python tools/sass_excerpt.py > profiles/rNN_sass_excerpts.txt"""
import collections, os, re, subprocess, sys
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(HERE, "calclens_b200", "libcalclens_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
print("# cuobjdump -sass calclens_b200/libcalclens_b200.so ; code objects:", ", ".join(arch))
want = [("legendre_synthesis_kernelILi3ELi2", "legendre_synthesis_kernel<3,2> (two shells per pass)", r"DFMA"),
        ("legendre_analysis_kernelILi8ELi2ELi2ELb0", "legendre_analysis_kernel<8,2,2,false> (two shells per pass)", r"DFMA"),
        ("legendre_synthesis_kernelILi4ELi1", "legendre_synthesis_kernel<4,1>", r"DFMA"),
        ("ring_synthesis_kernelE", "ring_synthesis_kernel", r"DADD|DMUL|DFMA"),
        ("ray_step_kernel", "ray_step_kernel", r"UBLKCP|SYNCS")]
funcs = re.split(r"\n\s*Function : ", txt)[1:]
for key, title, focus in want:
    f = [x for x in funcs if key in x.split("\n")[0]]
    if not f:
        print("\n## %s: not found" % title); continue
    f = f[0]
    ins = [m.group(2).strip() for m in (re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l) for l in f.split("\n")) if m]
    ops = collections.Counter()
    for t in ins:
        w = t.split()
        op = w[1] if w[0].startswith("@") and len(w) > 1 else w[0]
        ops[op.split(".")[0]] += 1
    print("\n## %s  (%s)" % (title, f.split("\n")[0][:60]))
    print("instructions %d; " % len(ins) + ", ".join("%s %d" % kv for kv in ops.most_common(14)))
    print("local-memory (spill) instructions: LDL %d, STL %d; cp.async LDGSTS %d; bulk-copy UBLKCP %d; mbarrier SYNCS %d; RED %d" % (
        ops["LDL"], ops["STL"], ops["LDGSTS"], ops["UBLKCP"], ops["SYNCS"], sum(v for k, v in ops.items() if k.startswith("RED"))))
    # densest 24-instruction window for the focus mnemonics
    hit = [1 if re.search(focus, t) else 0 for t in ins]
    W = 24
    best, bi = -1, 0
    s = sum(hit[:W])
    for i in range(0, max(1, len(ins) - W)):
        if s > best: best, bi = s, i
        s += (hit[i + W] if i + W < len(ins) else 0) - hit[i]
    print("excerpt (the %d consecutive instructions densest in %s):" % (W, focus))
    for t in ins[bi:bi + W]:
        print("    " + t)
