"""Pretty-print an HD_TRACE dump (tools/trace_step.py): python tools/show_trace.py trace.txt [t0_ms t1_ms]"""
import re, sys
rows, seen = [], False
for l in open(sys.argv[1]):
    if l.startswith("==== traced"):
        seen, rows = True, []
        continue
    m = re.match(r"\[hd_trace\] s(\d+)\s+([\d.\-]+)\s+([\d.\-]+)\s+([\d.\-]+) us\s+(\S+)", l)
    if m and seen:
        name = m.group(5)
        k = re.match(r"_ZN2hd\d+([a-z0-9_]+?)(?:E|I)", name)
        tmpl = re.findall(r"L[bi](\d+)E", name)
        rows.append((int(m.group(1)), float(m.group(2)), float(m.group(3)), (k.group(1) if k else name) + ("<" + ",".join(tmpl) + ">" if tmpl else "")))
rows.sort(key=lambda r: r[1])
t0 = rows[0][1]
lo, hi = (float(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (0, 1e9)
print(f"{len(rows)} launches, {rows[-1][2] - t0:.3f} ms")
busy = {}
for s, a, b, nm in rows:
    busy[s] = busy.get(s, 0) + (b - a)
    if lo <= a - t0 <= hi:
        print(f"{'        ' * s}s{s} {a - t0:8.3f} {b - t0:8.3f} {1e3 * (b - a):7.1f}  {nm}")
print({f"s{k}": round(v, 3) for k, v in busy.items()})
