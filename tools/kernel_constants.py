"""profiles/kernel_constants.json from `ncu --set full --import-source on` reports of the config-1 launch (3600 lags,
2048 x 2048 grid): executed FP64-pipe thread instructions per nominal pixel-sample (source page, predicated-on thread
instructions of the D* opcodes) and DRAM bytes per launch (raw page), stamped with the digest of the kernel sources
(`bench.csrc_digest`) so that `bench.py` quotes them only for the code they were measured on.

    python tools/kernel_constants.py "lag_corr_roll_kernel=gpurun_out/r2f_roll_fp64.ncu-rep" \
        "lag_corr_roll_kernel<MIXED>=gpurun_out/r2f_roll_mixed.ncu-rep" \
        "offset_window_kernel=gpurun_out/r2f_carr.ncu-rep@0.12987649*2048*2048*14400"
(`@expr`: the pixel-samples of that launch when they are not config 1's 2048*2048*3600 nominal ones; for the Carrington
kernel the evaluated (pixel, lag) pairs `tools/carr_lab.py` prints.)
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FP64_PIPE = ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def constants(rep, samples):
    raw = page(rep, "raw")
    val, unit = dict(zip(raw[0], raw[2])), dict(zip(raw[0], raw[1]))

    def dram(k):
        return float(val[k]) * UNIT[unit[k]]
    src = page(rep, "source")
    hdr = src[1]
    i_src, i_thr = hdr.index("Source"), hdr.index("Predicated-On Thread Instructions Executed")
    per_op = {}
    for row in src[2:]:
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", row[i_src])
        if m and row[i_thr].isdigit():
            per_op[m.group(1)] = per_op.get(m.group(1), 0) + int(row[i_thr])
    fp64 = sum(v for k, v in per_op.items() if k in FP64_PIPE)
    return {"kernel_name": val["Kernel Name"][:110], "gpu_time_ms": float(val["gpu__time_duration.sum"]),
            "fp64_instr_per_pixel_sample": fp64 / samples,
            "fp64_by_opcode_per_pixel_sample": {k: per_op[k] / samples for k in FP64_PIPE if k in per_op},
            "f2f_per_pixel_sample": per_op.get("F2F", 0) / samples,
            "thread_instr_per_pixel_sample": sum(per_op.values()) / samples,
            "dram_bytes_per_launch": dram("dram__bytes_read.sum") + dram("dram__bytes_write.sum"),
            "report": os.path.basename(rep)}


def main():
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+", help="name=path.ncu-rep")
    ap.add_argument("--samples", default="2048*2048*3600")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "kernel_constants.json"))
    args = ap.parse_args()
    samples = float(eval(args.samples, {}))     # arithmetic on the command line only
    out = {"csrc_sha256": bench.csrc_digest(),
           "what": "ncu --set full captures of one config-1 launch (3600 lags, 2048 x 2048 grid); per NOMINAL pixel-sample "
                   "(grid pixels x lags: about 3.5 % of them fall outside the small image and execute nothing)",
           "kernels": {}}
    for spec in args.reports:
        name, rep = spec.split("=", 1)
        n = samples
        if "@" in rep:
            rep, expr = rep.split("@", 1)
            n = float(eval(expr, {}))
        out["kernels"][name] = dict(constants(rep, n), pixel_samples_of_the_launch=n)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
