"""Executed instructions and stall samples of one kernel of an `ncu --set full --import-source on` report, grouped by the source
function each SASS instruction was inlined from (nvdisasm -gi line info of the SAME build of libdfb_b200.so).

  [SKIP=n] [LINES=m] python scripts/ncu_regions.py <report.ncu-rep> <kernel regex> <mangled-name fragment> [source.cu] > profiles/<name>.md
  (SKIP: launches of that kernel to skip in the report; LINES: rows of the per-line table)

The SASS page of the report and the disassembly of the cubin list the kernel's instructions in the same order, 16 bytes apart;
rows are joined by offset.  Function extents come from a scan of the sources for top-level definitions."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dynamicfusion_body_b200", "csrc")
LIB = os.path.join(ROOT, "dynamicfusion_body_b200", "libdfb_b200.so")


def function_extents():
    """{file basename: [(first line, last line, name)]} from a brace-depth scan (definitions at namespace depth)."""
    out = {}
    for fn in os.listdir(CSRC):
        if not fn.endswith((".cu", ".h")):
            continue
        spans, depth, cur, ns_depth = [], 0, None, 0
        lines = open(os.path.join(CSRC, fn)).read().split("\n")
        pending = None
        for i, line in enumerate(lines, 1):
            code = line.split("//")[0]
            if cur is None:
                if re.match(r"\s*namespace\b", code) or re.match(r'\s*extern "C"', code):
                    ns_depth += code.count("{")
                    continue
                m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*\(", code)
                if m and depth == 0 and not code.strip().startswith(("#", "}", "typedef", "using")) and m.group(1) not in ("if", "for", "while", "switch", "defined", "__launch_bounds__", "__align__", "sizeof"):
                    pending = (i, m.group(1)) if pending is None else pending
                if pending and "{" in code and depth == 0:
                    cur = pending
                    pending = None
                    depth = code.count("{") - code.count("}")
                    if depth == 0:
                        spans.append((cur[0], i, cur[1]))
                        cur = None
                    continue
                if pending and ";" in code and "{" not in code:
                    pending = None
                if code.strip() == "}" and ns_depth:
                    ns_depth -= 1
            else:
                depth += code.count("{") - code.count("}")
                if depth <= 0:
                    spans.append((cur[0], i, cur[1]))
                    cur, depth = None, 0
        out[fn] = spans
    return out


def disasm(fragment, source):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(source.split(".")[0] + ".")][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(txt) if l.startswith(".text.") and fragment in l]
    assert len(start) == 1, "kernel fragment matches %d functions" % len(start)
    rows, chain, fresh = [], [], True
    for l in txt[start[0] + 1:]:
        if l.startswith("//----") or l.startswith(".text."):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip(), list(chain)))
            fresh = True
    return rows


def main():
    rep, kregex, fragment = sys.argv[1:4]
    source = sys.argv[4] if len(sys.argv) > 4 else "tsdf.cu"
    ext = function_extents()

    def fn_of(loc):
        f, ln = loc
        for a, b, name in ext.get(f, []):
            if a <= ln <= b:
                return name
        return f if not f.endswith((".cu", ".h")) or f not in ext else "%s:%d" % (f, ln)

    dis = disasm(fragment, source)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kregex, "--launch-skip", os.environ.get("SKIP", "0"), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    lines = raw.split("\n")
    h = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[h:]))))
    seen, data = set(), []
    for r in rows:
        if r["Address"] in seen or not r["Address"].startswith("0x"):
            continue
        seen.add(r["Address"])
        data.append(r)
    base = int(data[0]["Address"], 16)
    assert len(data) == len(dis), "report has %d instructions, the library's kernel %d: not the same build" % (len(data), len(dis))
    stalls = [k for k in data[0] if k.startswith("stall_") and "Not Issued" not in k]
    tot_i = sum(int(r["Instructions Executed"]) for r in data)
    tot_s = sum(int(r["# Samples"]) for r in data)
    by_fn = collections.Counter(); by_fn_s = collections.Counter(); by_top = collections.Counter(); by_top_s = collections.Counter()
    for r, (off, text, chain) in zip(data, dis):
        assert int(r["Address"], 16) - base == off
        inner = fn_of(chain[0]) if chain else "?"
        # outermost frames first: the chain lists the innermost location, then its inline parents
        names = [fn_of(c) for c in chain]
        top = next((n for n in names if n in TOP_LEVEL), inner)
        n, s = int(r["Instructions Executed"]), int(r["# Samples"])
        by_fn[inner] += n; by_fn_s[inner] += s; by_top[top] += n; by_top_s[top] += s
    st = collections.Counter()
    for r in data:
        for k in stalls:
            st[k] += int(r[k])
    print("# %s: source-level view (ncu --set full --import-source on, one launch)\n" % kregex)
    print("warp instructions executed: %d; stall samples: %d\n" % (tot_i, tot_s))
    print("## stall reasons (all samples)\n\n| reason | samples | share |\n|---|---|---|")
    for k, v in st.most_common(9):
        print("| %s | %d | %.1f %% |" % (k, v, 100 * v / max(1, sum(st.values()))))
    print("\n## executed warp instructions by code region (the innermost of the kernel's stage functions each instruction was inlined through)\n")
    print("| region | warp instructions | share | stall samples |\n|---|---|---|---|")
    for k, v in by_top.most_common(14):
        print("| %s | %d | %.1f %% | %.1f %% |" % (k, v, 100 * v / tot_i, 100 * by_top_s[k] / max(1, tot_s)))
    print("\n## by innermost function\n\n| function | warp instructions | share | stall samples |\n|---|---|---|---|")
    for k, v in by_fn.most_common(22):
        print("| %s | %d | %.1f %% | %.1f %% |" % (k, v, 100 * v / tot_i, 100 * by_fn_s[k] / max(1, tot_s)))
    by_line = collections.Counter(); by_line_s = collections.Counter()
    for r, (off, text, chain) in zip(data, dis):
        if chain:
            by_line[chain[0]] += int(r["Instructions Executed"]); by_line_s[chain[0]] += int(r["# Samples"])
    print("\n## by source line (innermost location)\n\n| file:line | warp instructions | share | stall samples |\n|---|---|---|---|")
    for k, v in by_line.most_common(int(os.environ.get("LINES", "30"))):
        print("| %s:%d | %d | %.1f %% | %.1f %% |" % (k[0], k[1], v, 100 * v / tot_i, 100 * by_line_s[k] / max(1, tot_s)))
    print("\n## hottest instructions by stall samples\n\n| samples | share | executed | dominant stall | function | SASS |\n|---|---|---|---|---|---|")
    order = sorted(range(len(data)), key=lambda i: -int(data[i]["# Samples"]))[:25]
    for i in order:
        r = data[i]
        dom = max(stalls, key=lambda k: int(r[k]))
        chain = dis[i][2]
        print("| %s | %.2f %% | %s | %s | %s | `%s` |" % (r["# Samples"], 100 * int(r["# Samples"]) / max(1, tot_s), r["Instructions Executed"], dom,
                                                       fn_of(chain[0]) if chain else "?", dis[i][1]))
    op = collections.Counter()
    for r, d in zip(data, dis):
        t = d[1].split()
        name = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        op[name] += int(r["Instructions Executed"])
    print("\n## executed by opcode\n\n| opcode | warp instructions | share |\n|---|---|---|")
    for k, v in op.most_common(24):
        print("| %s | %d | %.1f %% |" % (k, v, 100 * v / tot_i))


TOP_LEVEL = {"stream_brick", "quad_pretest", "queue_process", "mixed_layer_quads", "mixed_layer", "brick_region_rec", "update_body",
             "proj_exact_kernel", "project_fuse_ref", "warp_ref", "dq_blend_ref", "normal_eq_data_kernel", "pcg_resident_kernel"}

if __name__ == "__main__":
    main()
