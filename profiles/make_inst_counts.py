"""profiles/inst_counts.json from ncu raw pages: warp-instructions per launch of the two kernels, normalised to the
unit bench.py scales them by (per 64-frame trellis step for the Viterbi kernel, per codeword for the RS kernel).

    python profiles/make_inst_counts.py profiles/ncu_r02/viterbi_pair_fic_raw.csv profiles/ncu_r02/viterbi_pair_msc_raw.csv \
        profiles/ncu_r02/rs_superframe_raw.csv

The raw pages come from `ncu --set full ... ; ncu -i rep --page raw --csv` (profiles/README.md).  Shapes are read from
the launch's own grid size and the file name tag; the FIC page is the bench workload (65,536 frames, F = 768)."""
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}  # bytes / microseconds


def page(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(d, key, default=None):
    v, u = d.get(key, ("", ""))
    try:
        return float(v.replace(",", "")) * SCALE.get(u, 1.0)
    except ValueError:
        return default


def main():
    fic, msc, rs = sys.argv[1:4]
    out = {}
    f = page(fic)
    inst = num(f, "smsp__inst_executed.sum")
    groups, steps = 65536 // 64, 774
    cyc = num(f, "smsp__cycles_active.avg")
    alu_pct = num(f, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
    alu = num(f, "smsp__inst_executed_pipe_alu.sum")
    if alu is None and alu_pct is not None and cyc:
        alu = alu_pct / 100.0 * 0.5 * cyc * 592  # the ALU pipe issues one warp-instruction every 2 cycles per sub-partition
    m = page(msc)
    m_inst, m_grid = num(m, "smsp__inst_executed.sum"), num(m, "launch__grid_size")
    m_groups, m_steps = 262144 // 64, 3078  # the MSC page is bench.py's extra.msc workload
    out["viterbi_pair_kernel"] = {
        "warp_inst_per_group_step": inst / (groups * steps),
        "alu_pipe_inst_per_group_step": alu / (groups * steps) if alu else None,
        "fmaheavy_pipe_inst_per_group_step": (num(f, "smsp__inst_executed_pipe_fmaheavy.sum") or 0) / (groups * steps) or None,
        "fmalite_pipe_inst_per_group_step": (num(f, "smsp__inst_executed_pipe_fmalite.sum") or 0) / (groups * steps) or None,
        "lsu_pipe_inst_per_group_step": (num(f, "smsp__inst_executed_pipe_lsu.sum") or 0) / (groups * steps) or None,
        "fic_launch": {"frames": 65536, "framebits": 768, "warp_inst": inst, "duration_us_under_ncu": num(f, "gpu__time_duration.sum"),
                       "registers": num(f, "launch__registers_per_thread"), "grid": num(f, "launch__grid_size"),
                       "issue_active_pct": num(f, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                       "alu_pipe_pct": alu_pct},
        "msc_launch": {"frames": 262144, "framebits": 3072, "warp_inst": m_inst, "grid": m_grid,
                       "warp_inst_per_group_step": m_inst / (m_groups * m_steps),
                       "alu_pipe_inst_per_group_step": (num(m, "smsp__inst_executed_pipe_alu.sum") or 0) / (m_groups * m_steps) or None,
                       "fmaheavy_pipe_inst_per_group_step": (num(m, "smsp__inst_executed_pipe_fmaheavy.sum") or 0) / (m_groups * m_steps) or None, "duration_us_under_ncu": num(m, "gpu__time_duration.sum"),
                       "registers": num(m, "launch__registers_per_thread"),
                       "issue_active_pct": num(m, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                       "alu_pipe_pct": num(m, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")},
        "dram_bytes_per_fic_launch": (num(f, "dram__bytes_read.sum", 0) + num(f, "dram__bytes_write.sum", 0)) or None,
        "source": "%s, %s" % (os.path.relpath(fic, os.path.dirname(HERE)), os.path.relpath(msc, os.path.dirname(HERE))),
    }
    r = page(rs)
    r_inst = num(r, "smsp__inst_executed.sum")
    tag = os.path.basename(rs)  # rs_superframe_s4_raw.csv -> 125,000 superframes of s = 4
    import re

    s = int(re.search(r"_s(\d+)_", tag).group(1))
    ncw = 125000 * s
    out["rs_superframe_kernel"] = {
        "warp_inst_per_codeword": r_inst / ncw, "launch": {"superframes": 125000, "rs_dims": s, "warp_inst": r_inst,
                                                          "duration_us_under_ncu": num(r, "gpu__time_duration.sum"),
                                                          "registers": num(r, "launch__registers_per_thread"),
                                                          "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")},
        "dram_bytes_per_codeword": ((num(r, "dram__bytes_read.sum", 0) + num(r, "dram__bytes_write.sum", 0)) / ncw) or None,
        "source": os.path.relpath(rs, os.path.dirname(HERE)),
    }
    with open(os.path.join(HERE, "inst_counts.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
