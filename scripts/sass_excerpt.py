"""SASS mnemonics of the production kernels -> profiles/<round>_sass_excerpt.md (cuobjdump -sass of the in-tree library).
usage: python scripts/sass_excerpt.py > profiles/r2_sass_excerpt.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dynamicfusion_body_b200", "libdfb_b200.so")
WANT = collections.OrderedDict([
    ("brick_update_kernelILi4ELb1ELb1", "brick_update_kernel<4, exact k, one view> (the 512^3 bench kernel)"),
    ("brick_update_smem_kernelILi8ELb1ELb0", "brick_update_smem_kernel<8, exact k, several views> (node table in shared memory, the 8-view k = 8 configuration)"),
    ("proj_exact_kernelILi4ELi4", "proj_exact_kernel<4, 4>"),
    ("region_bounds_kernel", "region_bounds_kernel"),
    ("normal_eq_data_kernelILi4", "normal_eq_data_kernel<4>"),
    ("pcg_pipelined_kernelILi1", "pcg_pipelined_kernel<1>"),
])
PICK = r"(UBLKCP|SYNCS|LDS\.128|LDS\.64|LDG\.E\.EF|STG\.E\.EF|LDG\.E\.128|CCTL|RED|ATOM|DFMA|DADD|DMUL|MUFU|SHFL|BAR|MEMBAR|UTMA|HMMA|UTC)"


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    pat = re.compile(r"/\*[0-9a-f]{4,}\*/\s+(.*?);")
    out = ["# round 2: SASS mnemonics of the production kernels (cuobjdump -sass libdfb_b200.so, sm_100a)\n",
           "What to look for: `UBLKCP` = TMA bulk copy (cp.async.bulk) of the packed node table into shared memory, `SYNCS` = its mbarrier; `LDS.128` = node records read",
           "from that table; `LDG.E.EF` / `STG.E.EF` (evict-first) = the streaming reads / writes of voxel values and weights; `LDG.E.128` = node records / region records",
           "from global memory; `CCTL.E.PF1` = the L1 prefetch at queue push; `RED.E.ADD.F64` = block accumulation of the normal equations; `DFMA` = the float64 exact tier",
           "and the PCG.  No `UTC*MMA` / `HMMA` / `UTMALDG` anywhere: the path has no dense contraction to put on the tensor pipes (DESIGN.md section 4).\n"]
    allops = collections.Counter()
    found = {}
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        ops = collections.Counter()
        for m in pat.finditer(f):
            t = m.group(1).split()
            ops[t[1] if t[0].startswith("@") else t[0]] += 1
        for o, c in ops.items():
            allops[o.split(".")[0]] += c
        for k in WANT:
            if k in name and k not in found:
                found[k] = (name, ops, f)
    for k, title in WANT.items():
        if k not in found:
            continue
        name, ops, f = found[k]
        out.append("## %s\n\n`%s`: %d instructions\n" % (title, name[:110], sum(ops.values())))
        out.append("| mnemonic | static count |\n|---|---|")
        for o, c in sorted(((o, c) for o, c in ops.items() if re.match(PICK, o)), key=lambda kv: -kv[1])[:28]:
            out.append("| `%s` | %d |" % (o, c))
        lines = f.split("\n")
        for needle in ("UBLKCP", "LDG.E.EF", "CCTL.E.PF1", "RED.E.ADD.F64"):
            idx = [i for i, l in enumerate(lines) if needle in l]
            if idx:
                i = idx[0]
                ex = [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).rstrip() for l in lines[max(0, i - 2):i + 3] if "/*" in l and not l.strip().startswith("/* 0x")]
                out.append("\n```\n" + "\n".join(e[:140] for e in ex) + "\n```")
        out.append("")
    out.append("## whole library: tensor / TMA-tensor mnemonics\n")
    for o in ("UTCHMMA", "UTCQMMA", "UTCMMA", "HMMA", "IMMA", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS"):
        out.append("- `%s`: %d" % (o, sum(c for k, c in allops.items() if k.startswith(o))))
    print("\n".join(out))


if __name__ == "__main__":
    main()
