"""Turn the ncu captures of one profiling run (gpurun_out/) into the committed summaries under profiles/:
   launch list (md + csv), per-launch tables of the --set full captures, and roofline_traffic.json (DRAM bytes / launch)."""
import csv, json, os, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
B = 128
MACS = {"conv1_2": 603.98e6, "conv2_1": 150.99e6, "conv2_2": 150.99e6, "conv3_1": 75.50e6, "conv3_2": 150.99e6,
        "conv4_1": 37.75e6, "conv4_2": 37.75e6, "conv1_1": 18.87e6}


NOTES_CONV = [
    "Reading.  conv1_2.fwd runs on ROW TILES fused with pool1 (`tc_conv_kernel<64,1,2,1,0,1>`): DRAM read 277 MB = the bf16 input stream once, DRAM write 64 MB = the pooled",
    "stream + routing codes (the full-resolution activation is never written), tensor pipe active 67 % — its MMAs are organised by input row so that half of them are",
    "N = 128 (two adjacent weight blocks -> two adjacent accumulators).  An (M128,N64,K16) MMA cannot go below ~50 clk (A-operand fetch, tools/tc_probe2.py,",
    "profiles/r1_tc_probe2.log: 49.8 clk with cta_group::1 AND cta_group::2), i.e. plain N = 64 layers cap at 64 % of the tensor peak; the stream-tiled 64-channel launches",
    "(conv2_x) sit at 46-49 % active = 72-76 % of that cap with 15 items per CTA.  conv1_2.dgrad (`<64,1,2,1,1,2>`, row tiles, TMA-store epilogue + 1-bit ReLU mask) is",
    "EPILOGUE-paced (PC sampling of the previous build: the epilogue warps spent 5 % of their time waiting for accumulators); with two epilogue warps per TMEM lane quarter",
    "(384 threads, 111-124 registers) it went from 169 us / 47 % to 152 us / 53 % tensor-pipe active.  The <128,2,2,0,*> launches are the 32x32 / 16x16 layers: few tiles",
    "(162 items for 148 SMs at 16x16) and weights streamed from L2 (288 KB per item), hence the lower utilisation; they are 17 % of the family's time.",
]
NOTES_WG = [
    "Reading.  tc_wgrad64_kernel (64 -> 64 layers) issues TWO (M128,N192,K16) MMAs per 16 positions (A = two row shifts of X, B = three row shifts of G): conv1_2.wgrad",
    "reaches 86.9 % tensor-pipe active / 1190 TFLOP/s algorithmic (the previous five-MMA N = 64 formulation was capped at 64 %); DRAM read 554 MB = X and G once each.",
    "The capture starts mid-step (conv1_tc_wgrad_kernel also matches the regex), so the rows are conv2_1, conv1_2, conv1_1 of one step and conv4_2..conv3_1 of the next.",
    "conv1_tc_wgrad_kernel (K = 18) is bound by reading the 277 MB gradient stream (54 % of DRAM peak, 3 CTAs / SM).  The 16x16 / 32x32 layers are small (21-76 MB,",
    "26-54 % active): accumulators stay in TMEM for the CTA's whole run and are reduced with 16-byte vector REDs.",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[2:]


def table(rep, labels, title, cmd, notes):
    h, data = raw(rep)
    g = lambda r, n: r[h.index(n)]
    f = lambda r, n: float(g(r, n).replace(",", ""))
    lines = ["# " + title, "", "Command: `" + cmd + "` (one B200, under gpurun).",
             "Times are under the profiler (cold cache, serialised): compare SHARES and counters; bench.py reports the live CUDA-event times (`phases_ms_per_step`).", "",
             "| launch | kernel | time us | DRAM rd MB | DRAM wr MB | L2->SM MB | tensor-pipe active % | DRAM % of peak | TFLOP/s (algorithmic) | grid | regs |",
             "|---|---|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    tot_t = tot_f = tot_b = 0.0
    for r, lab in zip(data, labels):
        name = g(r, "Kernel Name").split("(")[0].replace("void ", "").replace("udh::tc::", "").replace("tc::", "")
        t = f(r, "gpu__time_duration.sum")
        unit_t = 1.0          # ncu raw page reports us here
        rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
        l2 = f(r, "l1tex__m_xbar2l1tex_read_bytes.sum")
        tp = f(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")
        dp = f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        layer = lab.split(".")[0]
        flops = 2.0 * MACS.get(layer, 0.0) * B
        tf = flops / (t * 1e-6) / 1e12 if flops else 0.0
        lines.append("| %s | %s | %.1f | %.1f | %.1f | %.1f | %.2f | %.2f | %.0f | %s | %s |" % (
            lab, name, t, rd, wr, l2, tp, dp, tf, g(r, "launch__grid_size"), g(r, "launch__registers_per_thread")))
        traffic[lab] = (rd + wr) * 1e6
        tot_t += t; tot_f += flops; tot_b += (rd + wr)
    lines += ["", "Family total: %.1f us, %.1f GFLOP algorithmic -> %.0f TFLOP/s; DRAM traffic %.0f MB over these launches." % (
        tot_t, tot_f / 1e9, tot_f / (tot_t * 1e-6) / 1e12, tot_b), ""] + notes
    return "\n".join(lines) + "\n", traffic


def units_check(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    return {n: u[h.index(n)] for n in ("gpu__time_duration.sum", "dram__bytes_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum")}


def launches(csv_path, cmd):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").replace("udh::", "").replace("tc::", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += float(r[-1]) / 1e3
    tot = sum(v[1] for v in agg.values())
    lines = ["# ncu launch list, bf16 mode, B=128 (round 1, final kernels)", "", "Command: `" + cmd + "`",
             "(%d consecutive launches; per-launch times are cold-cache and serialised: compare SHARES with bench.py's `phases_ms_per_step`)." % len(rows), "",
             "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| %s | %d | %.1f | %.1f %% |" % (k, c, t, 100 * t / tot))
    lines += ["", "total %.0f us over %d launches" % (tot, len(rows))]
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    go = os.path.join(ROOT, "gpurun_out")
    print(units_check(os.path.join(go, "prof_conv_r1c.ncu-rep")))
    bench = "python bench.py --steps 1 --warmup 3 --numeric bf16 --no-e2e --no-cpu-baseline --no-extras"
    conv_labels = ["conv1_2.fwd", "conv2_1.fwd", "conv2_2.fwd", "conv3_1.fwd", "conv3_2.fwd", "conv4_1.fwd", "conv4_2.fwd",
                   "conv4_2.dgrad", "conv4_1.dgrad", "conv3_2.dgrad", "conv3_1.dgrad", "conv2_2.dgrad", "conv2_1.dgrad", "conv1_2.dgrad"]
    md, tr_conv = table(os.path.join(go, "prof_conv_r1c.ncu-rep"), conv_labels,
                        "ncu --set full — tc_conv_kernel (tcgen05 implicit-GEMM conv, fwd + dgrad), one train step (B=128, bf16 mode), round 1 (final kernels)",
                        "ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 42 -c 14 " + bench, NOTES_CONV)
    open(os.path.join(OUT, "r1_ncu_tc_conv.md"), "w").write(md)
    wg_labels = ["conv2_1.wgrad", "conv1_2.wgrad", "conv1_1.wgrad", "conv4_2.wgrad", "conv4_1.wgrad", "conv3_2.wgrad", "conv3_1.wgrad"]
    md, tr_wg = table(os.path.join(go, "prof_wgrad_r1c.ncu-rep"), wg_labels,
                      "ncu --set full — weight-gradient kernels (tc_wgrad64_kernel N=192, tc_wgrad_kernel, conv1_tc_wgrad_kernel), B=128, bf16 mode, round 1 (final kernels)",
                      "ncu --set full --clock-control none --import-source on -k regex:tc_wgrad -s 21 -c 7 " + bench, NOTES_WG)
    open(os.path.join(OUT, "r1_ncu_tc_wgrad.md"), "w").write(md)
    lcmd = "ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 170 --csv python bench.py --steps 2 --warmup 3 --numeric bf16 --no-e2e --no-cpu-baseline --no-extras"
    open(os.path.join(OUT, "r1_launches_bf16.md"), "w").write(launches(os.path.join(go, "launches_r1c.csv"), lcmd))
    with open(os.path.join(OUT, "r1_launches_bf16.csv"), "w") as f:
        for line in open(os.path.join(go, "launches_r1c.csv")):
            if line.startswith('"'):
                f.write(line)
    tp = os.path.join(OUT, "roofline_traffic.json")
    tr = json.load(open(tp)) if os.path.exists(tp) else {}
    tr["tc_conv_kernel"] = sum(tr_conv.values()) / len(tr_conv)
    wg_only = [v for k, v in tr_wg.items() if k != "conv1_1.wgrad"]
    tr["tc_wgrad_kernel"] = sum(wg_only) / len(wg_only)
    tr.setdefault("per_launch_bytes", {}).update(tr_conv); tr["per_launch_bytes"].update(tr_wg)
    tr["per_launch_bytes"].pop("conv2_2.wgrad?", None)
    tr["source"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full captures of the final round-1 kernels (profiles/r1_ncu_tc_conv.md, r1_ncu_tc_wgrad.md); family entries are the mean over the family's launches of one step"
    json.dump(tr, open(tp, "w"), indent=1)
    print("written")
