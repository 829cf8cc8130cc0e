"""In-order single-warp issue model of a kernel's hottest FP64 window, from cuobjdump SASS (no GPU needed).

   python tools/sass_stall_model.py <object or .so> <mangled kernel name> [window]

Finds the `window` consecutive instructions holding the most DFMA/DADD/DMUL and steps through them the way one
warp alone would issue them: in order, one FP64 instruction per 2 cycles, a result usable 8 cycles after issue
(measured on sm_100a: tools/ubench/dfma_lat.cu), LDS 30, SHFL 25, everything else 4.  The stall cycles it
reports are the issue slots a co-resident warp has to fill; it is a schedule-quality indicator, not a timing."""
import re
import subprocess
import sys


def main():
    obj, fun = sys.argv[1], sys.argv[2]
    win = int(sys.argv[3]) if len(sys.argv) > 3 else 540
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
    ins = [m.group(1) for m in (re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", l) for l in txt.splitlines()) if m]
    fp = lambda o: any(x in o for x in ("DFMA", "DADD", "DMUL"))
    best = max(range(max(1, len(ins) - win)), key=lambda i: sum(1 for o in ins[i:i + win] if fp(o)))
    loop = ins[best:best + win]
    regs = lambda t: [int(x) for x in re.findall(r"\bR(\d+)", t)]
    ready, cyc, stall_total, hist = {}, 0, 0, {}
    for op in loop:
        body = op.split(" ", 1)[1] if op.startswith("@") else op
        name = body.split()[0]
        ops = body[len(name):].split(",")
        dst, srcs = (ops[0], ops[1:]) if not name.startswith("STS") else ("", ops)
        need = max([ready.get(r + q, 0) for s in srcs for r in regs(s) for q in (0, 1)] + [0])
        t = max(cyc, need)
        if t > cyc:
            hist[name.split(".")[0]] = hist.get(name.split(".")[0], 0) + t - cyc
            stall_total += t - cyc
        isfp = fp(name)
        lat = 8 if isfp else 30 if name.startswith("LDS") else 25 if name.startswith("SHFL") else 4
        width = 2 if (isfp or ".64" in name) else 4 if ".128" in name else 1
        for r in regs(dst):
            for q in range(width):
                ready[r + q] = t + lat
        cyc = t + (2 if isfp else 1)
    nfp = sum(1 for o in loop if fp(o))
    print(f"window {best}..{best + win}: {nfp} FP64 instructions = {2 * nfp} pipe cycles; one warp alone needs {cyc} "
          f"cycles ({stall_total} stalled: {hist})")


if __name__ == "__main__":
    main()
