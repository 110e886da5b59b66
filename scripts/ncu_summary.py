"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.txt [note]"""
import csv
import collections
import io
import subprocess
import sys

KEYS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum dram__throughput.avg.pct_of_peak_sustained_elapsed
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
launch__registers_per_thread launch__grid_size launch__block_size launch__shared_mem_per_block_dynamic launch__waves_per_multiprocessor
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem launch__occupancy_limit_warps
sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.sum sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.sum sm__inst_executed_pipe_lsu.sum sm__inst_executed_pipe_xu.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct lts__t_bytes.sum sm__cycles_elapsed.max smsp__cycles_active.avg
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active""".split()


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of {rep}", f"# {note}", ""]
    for k, vals in enumerate(rows[2:]):
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"## launch {k}: {name}")
        for i, h in enumerate(hdr):
            if h in KEYS:
                lines.append(f"{h:72s} {units[i]:16s} {vals[i]}")
        lines.append("-- warp stall reasons (pct of warp-active cycles, > 2 %)")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_warp_active.pct"):
                try:
                    v = float(vals[i])
                except ValueError:
                    continue
                if v > 2:
                    lines.append(f"{h:72s} {v:8.2f}")
        lines.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 3:
        body = [r for r in srows[2:] if len(r) > 6 and r[5].isdigit()]
        tot = sum(int(r[5]) for r in body)
        op = collections.Counter()
        for r in body:
            s = r[1].strip()
            if s.startswith("@"):
                s = s.split(None, 1)[1]
            op[s.split()[0]] += int(r[5])
        lines.append(f"-- executed warp instructions by opcode (first kernel, total {tot}, {len(body)} SASS lines)")
        for o, c in op.most_common(22):
            lines.append(f"{o:32s} {c:12d} {100 * c / tot:5.1f}%")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
