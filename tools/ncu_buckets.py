"""Summarise an .ncu-rep (first kernel): key raw metrics, instruction buckets by execution
count, top stall lines.  Usage: python tools/ncu_buckets.py gpurun_out/x.ncu-rep [ntop]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum', 'lts__t_bytes.sum']
for h, u, v in zip(hdr, units, vals):
    if h in want or ('stalled' in h and 'per_issue_active' in h and 'not_issued' not in h and float(v or 0) > 0.05):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
num = lambda x: float(x) if re.match(r'^-?[\d.]+$', x or '') else 0.0
b = collections.defaultdict(lambda: [0, 0, 0])
for r in data:
    n = num(r[ix['Instructions Executed']]); s = num(r[ix['# Samples']])
    b[n][0] += 1; b[n][1] += n; b[n][2] += s
tot = sum(v[1] for v in b.values()); tots = sum(v[2] for v in b.values())
print(f"total instr {tot/1e6:.1f}M samples {tots:.0f}")
for n, v in sorted(b.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  exec/instr {n:10.0f} ({n/256:8.1f}/utt) #instr {v[0]:5d} total {v[1]/1e6:7.2f}M ({100*v[1]/tot:4.1f}%) samples {v[2]:6.0f} ({100*v[2]/tots:4.1f}%)")
top = sorted(data, key=lambda r: -num(r[ix['# Samples']]))[:ntop]
for r in top:
    st = {k: num(r[ix[k]]) for k in ('stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_mio', 'stall_barrier', 'stall_math', 'stall_not_selected', 'stall_branch_resolving', 'stall_dispatch', 'stall_lg', 'stall_no_inst', 'stall_sleep', 'stall_misc')}
    main = max(st, key=st.get)
    print(f"{num(r[ix['# Samples']]):6.0f} x{num(r[ix['Instructions Executed']]):9.0f} {main[6:]:>10s} | {r[ix['Source']].strip()[:80]}")
