"""Opcode histograms and the hot loops of the shipped kernels, from the built library (cuobjdump -sass; no GPU needed).
Writes profiles/<tag>_sass_<kernel>.txt: what proves the Blackwell-native claims (UTMALDG / SYNCS for TMA + mbarrier, FFMA2 packed
fp32x2, LDG.E.128 of the z-quad kernels) and how many instructions a sample costs."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tomography_alignment_b200", "libtomo_b200.so")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
KERNELS = ["ray_kernel_forward", "ray_kernel_gradient", "adjoint_tile_kernel", "voxel_bilinear_tma_kernel", "sep_forward_kernel",
           "sep_adjoint_kernel", "sep_gradient_kernel", "zq_kernel_forward", "zq_kernel_gradient"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs, cur = {}, None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln):
        funcs[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln).rstrip())


def opcode(ln):
    m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    return m.group(1) if m else None


def loops(lines):
    """(start, end) of backward branches: innermost loops."""
    addr = {int(re.search(r"/\*([0-9a-f]{4,6})\*/", l).group(1), 16): i for i, l in enumerate(lines)}
    out = []
    for i, l in enumerate(lines):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s+)?0x([0-9a-f]+)", l)
        if m:
            t = int(m.group(1), 16)
            if t in addr and addr[t] <= i:
                out.append((addr[t], i))
    return out


for k in KERNELS:
    name = [f for f in funcs if k in f]
    if not name:
        continue
    lines = funcs[name[0]]
    hist = collections.Counter(opcode(l).split(".")[0] for l in lines if opcode(l))
    full = collections.Counter(opcode(l) for l in lines if opcode(l))
    lp = loops(lines)
    # the hot loop: the SHORTEST backward-branch span holding at least MIN_MEM[k] memory instructions (the sample loop of the
    # tile kernel contains a rarely taken 16-access sub-loop; the TMA kernel's loop body is 32 unrolled voxels)
    def nmem(se):
        return sum(1 for l in lines[se[0]:se[1] + 1] if re.search(r"\b(LDG|LDS|STS|LDGSTS)\b", l.replace(".", " ")))
    need = {"adjoint_tile_kernel": 17, "voxel_bilinear_tma_kernel": 100, "ray_kernel_forward": 16}.get(k, 8 if "ray" in k or "zq" in k else 4)
    cand = [se for se in lp if nmem(se) >= need]
    hot = min(cand, key=lambda se: se[1] - se[0]) if cand else None
    path = os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (TAG, k))
    with open(path, "w") as f:
        f.write("# %s (%s), %d SASS instructions; cuobjdump -sass of tomography_alignment_b200/libtomo_b200.so (sm_100a)\n" % (k, name[0], len(lines)))
        f.write("# opcode histogram (static):\n")
        for op, c in hist.most_common():
            f.write("#   %-10s %d\n" % (op, c))
        marks = [op for op in full if re.match(r"(UTMALDG|UTMASTG|UBLKCP|SYNCS|FFMA2|FMUL2|FADD2|LDG\.E\.128|LDS\.128|REDUX|ATOMS|DFMA|DADD)", op)]
        f.write("# notable: %s\n" % ", ".join("%s x%d" % (op, full[op]) for op in sorted(marks)))
        if hot:
            body = lines[hot[0]:hot[1] + 1]
            f.write("# hot loop: %d instructions (%s)\n" % (len(body), ", ".join("%s x%d" % (o, c) for o, c in collections.Counter(opcode(l).split(".")[0] for l in body if opcode(l)).most_common(12))))
            f.write("\n".join(body) + "\n")
    print(k, len(lines), "hot loop", (hot[1] - hot[0] + 1) if hot else None)
