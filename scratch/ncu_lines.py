"""Per-source-line instruction / stall-sample table from an ncu report (cuda,sass source page).
usage: python scratch/ncu_lines.py REPORT.ncu-rep KERNEL_INDEX [TOP]"""
import csv, collections, subprocess, sys, io
rep, idx = sys.argv[1], int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:stream_kernel",
                      "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; agg = collections.OrderedDict(); tot = tots = 0; kname = ""
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': kname = r[1]; continue
    if r[0] == 'Line No': continue
    if r[0] != '':
        try: n = int(r[7]); s = int(r[6])
        except Exception: continue
        a = agg.setdefault((cur, int(r[0])), [0, 0, r[1][:110]]); a[0] += n; a[1] += s; tot += n; tots += s
print(kname[:100]); print('total warp-instr', tot, 'samples', tots)
for (f, l), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{f[:20]:20s}:{l:4d} {100*n/tot:5.1f}% smp {100*s/max(tots,1):5.1f}%  {src}')
