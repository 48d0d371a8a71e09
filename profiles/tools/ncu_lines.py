#!/usr/bin/env python
"""Per-source-line instruction / stall summary of one kernel in an .ncu-rep (needs --import-source on, -lineinfo).
usage: ncu_lines.py report.ncu-rep kernel-substring [top N]"""
import sys, collections
sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ctx = ncu_report.load_report(rep)
for ri in range(ctx.num_ranges()):
    rng = ctx.range_by_idx(ri)
    for ai in range(rng.num_actions()):
        act = rng.action_by_idx(ai)
        if pat not in act.name():
            continue
        inst = act.metric_by_name("inst_executed")
        samp = act.metric_by_name("smsp__pcsamp_sample_buffer") or None
        pcs = act.metric_by_name("smsp__pcsamp_warps_issue_stalled_long_scoreboard")
        n = inst.num_instances()
        cor = inst.correlation_ids()
        by_line = collections.Counter(); by_line_s = collections.Counter()
        tot = 0
        stall_names = [m for m in act.metric_names() if m.startswith("smsp__pcsamp_warps_issue_stalled_") and not m.endswith("_not_issued")]
        stall_tot = collections.Counter()
        for i in range(n):
            pc = cor.as_uint64(i)
            v = inst.as_uint64(i)
            tot += v
            si = act.source_info(pc)
            key = (si.file_name().split("/")[-1], si.line()) if si else ("?", 0)
            by_line[key] += v
        for sn in stall_names:
            mm = act.metric_by_name(sn)
            cc = mm.correlation_ids()
            for i in range(mm.num_instances()):
                v = mm.as_uint64(i)
                if not v: continue
                si = act.source_info(cc.as_uint64(i))
                key = (si.file_name().split("/")[-1], si.line()) if si else ("?", 0)
                by_line_s[key] += v
                stall_tot[sn.replace("smsp__pcsamp_warps_issue_stalled_", "")] += v
        stot = sum(by_line_s.values())
        print(f"== {act.name()}  inst_executed {tot}  samples {stot}")
        print("   stalls:", ", ".join(f"{k} {100*v/stot:.1f}%" for k, v in stall_tot.most_common(8)))
        for (f, l), v in by_line.most_common(top):
            print(f"   {f}:{l:<5} inst {100*v/tot:5.1f}%   stall {100*by_line_s[(f,l)]/max(stot,1):5.1f}%")
        sys.exit(0)
