"""Turn the ncu artefacts in gpurun_out/ into the committed summaries under profiles/.
usage: python tools/summarize_ncu.py r01"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)


def short(name):
    m = re.search(r"conv_igemm_kernel<([^>]*)>", name)
    if m:
        a = [re.sub(r"\((?:int|bool)\)", "", x).strip() for x in m.group(1).split(",")]
        a += ["0"] * (7 - len(a))
        epi = {"0": "P8", "1": "PIXSHUF", "2": "FINAL"}.get(a[2], a[2])
        flags = ("" if a[4] == "0" else ",X3") + ("" if a[5] == "0" else ",PAIR") + ("" if a[6] == "0" else ",QUAD")
        return f"conv_igemm<NT={a[0]},KS={a[1]},{epi},SX={a[3]}{flags}>"
    m = re.search(r"rdb_tail_kernel<([^>]*)>", name)
    if m:
        a = [re.sub(r"\((?:int|bool)\)", "", x).strip() for x in m.group(1).split(",")] + ["0"]
        return "rdb_tail_kernel<" + ("streams" if a[0] == "1" else "handoff") + (",quadt" if a[1] == "1" else "") + ">"
    return re.sub(r"\(.*", "", name).replace("binb::", "").replace("void ", "")


def launches(csvname="launches_window.csv", outname=None, cmd_note=None, title="ncu launch list of ONE steady-state 6-frame 1280x720 window"):
    path = os.path.join(G, csvname)
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else v * 1e3 if r["Metric Unit"] == "ms" else v
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(os.path.join(P, outname or f"{tag}_launches_window.md"), "w") as f:
        f.write(f"# {tag}: {title}\n\n" +
                (cmd_note or (f"Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s {528 + len(rows)} -c {len(rows)} --csv python tools/run_window.py 2`\n"
                f"(BIN_B200_GRAPH=0; skip = 528 weight-pack launches + the {len(rows)} launches of window 0).")) + " Per-launch times under ncu are\n"
                "cold-cache and serialised: compare SHARES, not absolutes.\n\n"
                f"launches: {len(rows)}, sum of kernel durations: {tot/1e3:.2f} ms\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {t:.1f} | {t/n:.1f} | {100*t/tot:.1f}% |\n")
    print("wrote launches summary", len(rows), "launches")


WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor(hmma) inst % of peak"),
        ("TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "hmma cycles active (per TPC, 2 SMs)"),
        ("sm__cycles_elapsed.max", "SM cycles elapsed"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]


def full(rep, title, note, extra=()):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        return
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    base = re.sub(r"^r\d+[a-z]\d*_", "", rep.replace(".ncu-rep", ""))          # r02f_prof_tail -> prof_tail
    with open(os.path.join(P, f"{tag}_{base}.md"), "w") as f:
        f.write(f"# {tag}: {title}\n\n{note}\n\n| metric | " + " | ".join(short(d[ix['Kernel Name']]) for d in data) + " |\n")
        f.write("|---|" + "---|" * len(data) + "\n")
        for m, label in list(WANT) + list(extra):
            if m in ix:
                f.write(f"| {label} [{units[ix[m]]}] | " + " | ".join(d[ix[m]][:12] for d in data) + " |\n")
    print("wrote", rep)


if tag == "r01":
    launches()
    full("prof_rdb5.ncu-rep", "ncu --set full: RDB 5 of the stage-1 launch (5 batched calls, 360x640): conv0..conv3 (x-stacked) + LFF",
         "Command: `ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 357 -c 5 python tools/run_window.py 2`.\n"
         "Algorithmic bytes per launch: conv c reads 5*230400*(192+64c) B and writes 5*230400*64 B; LFF reads 5*230400*640 B, writes 5*230400*192 B.\n"
         "Algorithmic FLOPs per launch: 2*5*230400*(96+32c)*32*9 (conv c), 2*5*230400*224*96 (LFF).")
    full("prof_wgrad.ncu-rep", "ncu --set full: weight-gradient GEMM (tcgen05 MN-major) inside a training step, batch 4 x 256x256",
         "Command: `BT_STEPS=1 BT_WARM=1 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 300 -c 2 python tools/bench_train.py 4 256 256` (captured before the slab-reduce flush replaced the atomics).")
    full("prof_tail.ncu-rep", "ncu --set full: tail of the same stage: GFF.0, GFF.1, UPNet.0(+PixelShuffle), UPNet.2(+mean)",
         "Command: `ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 392 -c 4 python tools/run_window.py 2`.")
    full("prof_rdb_tail.ncu-rep", "ncu --set full: fused RDB tail (conv3 3x3+ReLU, LFF 1x1, residual) of the stage-1 launch (5 batched calls, 360x640)",
         "Command: `ncu --set full --clock-control none --import-source on -k regex:rdb_tail_kernel -s 48 -c 2 python tools/run_window.py 2`.\n"
         "Algorithmic bytes per launch: reads 5*230400*384 B (x + g0..g2) + 5*230400*192 B (residual, L2-resident), writes 5*230400*192 B.\n"
         "Algorithmic FLOPs per launch: 2*5*230400*(192*32*9 + 224*96).")

if tag == "r02":
    launches("r02t_launches_train.csv", "r02_launches_train_step.md",
             "Command: `BT_STEPS=1 BT_WARM=1 BIN_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/bench_train.py 4 256 256` "
             "(2 steps: fwd+bwd+loss+Adam, 6-frame net, batch 4 x 256x256; final tree).",
             title="ncu launch list of a training run (2 optimize_parameters steps)")
    launches("r02z_launches_window.csv", "r02_launches_window.md",
             "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 227 -c 223 --csv python tools/run_window.py 2` "
             "(BIN_B200_GRAPH=0; skip = 4 batched weight-pack launches + the 223 launches of window 0; default switches: four MMA warps, single-CTA kernels; final tree of round 2).")
    PORT = [("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem data pipe: tensor-core operand reads % of peak"),
            ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem data pipe: LSU (SHFL / LDS / mbarrier) % of peak"),
            ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed", "  of which shared loads % of peak")]
    FP = [("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak"),
          ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
          ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
    full("r02f_prof_k345.ncu-rep", "ncu --set full: the memory-/FP32-bound kernels of one window: K3 pack_frames, K5 convlstm, (K4 = final conv, see r02_prof_final)",
         "Command: `ncu --set full --clock-control none --import-source on -k regex:'pack_frames_kernel|convlstm_kernel|conv_igemm_kernel<16' -s 5 -c 5 python tools/run_window.py 2` (K3 stage-1 launch, K5 3-cell launch, K3, K4 final conv, K3).\n"
         "pack_frames: reads 3n fp32 frames (24n B per low-res position; frames shared by adjacent calls hit L2), writes 16 B per 8-channel plane.\n"
         "convlstm (prev_state=None): 324 FMA + 15 transcendentals per pixel against 36 B -> FP32-FMA bound, not HBM bound (DESIGN 4d).", FP)
    full("r02m_prof_conv_quad.ncu-rep", "ncu --set full: x-stacked RDB convs 0..2 at 5x360x640, final state (four MMA warps, 448 threads)",
         "Command: `ncu --set full --clock-control none --import-source on -k regex:'conv_igemm_kernel<\\(int\\)32' -s 144 -c 3 python tools/run_window.py 2`.\n"
         "Algorithmic FLOPs per launch: 2*5*230400*(96+32c)*32*9; bytes: reads 5*230400*(192+64c), writes 5*230400*64.\n"
         "The two `smem data pipe` rows add up to 90 / 92 / 95 % of the shared-memory data pipe's peak: the kernel is bound by that pipe (DESIGN 4a).", PORT)
    full("r02f_prof_tail.ncu-rep", "ncu --set full: fused RDB tail (default single-CTA kernel) at 5x360x640, final state of round 2",
         "Command: `ncu --set full --clock-control none --import-source on -k regex:rdb_tail -s 48 -c 1 python tools/run_window.py 2`.", PORT)
    for pr in ("0", "1"):
        full(f"r02f_prof_conv_pair{pr}.ncu-rep", f"ncu --set full: x-stacked RDB convs 0..2 at 5x360x640, BIN_B200_PAIR={pr} ({'CTA-pair cta_group::2' if pr == '1' else 'single-CTA'} kernel), same box",
             f"Command: `BIN_B200_PAIR={pr} ncu --set full --clock-control none --import-source on -k regex:'conv_igemm_kernel<\\(int\\)32' -s 144 -c 3 python tools/run_window.py 2`.\n"
             "Algorithmic FLOPs per launch: 2*5*230400*(96+32c)*32*9; bytes: reads 5*230400*(192+64c), writes 5*230400*64.")
    for pr in ("0", "1"):
        full(f"r02b_prof_tail_pair{pr}.ncu-rep", f"ncu --set full: fused RDB tail at 5x360x640, BIN_B200_PAIR={pr} ({'CTA-pair cta_group::2' if pr == '1' else 'single-CTA'} kernel), same box",
             f"Command: `BIN_B200_PAIR={pr} ncu --set full --clock-control none --import-source on -k regex:rdb_tail -s 48 -c 1 python tools/run_window.py 2`.\n"
             "Algorithmic bytes per launch: reads 5*230400*384 B (x + g0..g2) + residual (L2), writes 5*230400*192 B; FLOPs 2*5*230400*(192*32*9 + 224*96).")
