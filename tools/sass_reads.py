"""Register-file reads per FP64 instruction in a kernel's hottest window, from cuobjdump SASS (no GPU needed).
A DFMA that reads three distinct 64-bit registers issues every ~2.8 cycles instead of 2 on sm_100a
(tools/ubench/dfma_operands.cu: 26.0 vs 36.8 TFLOP/s); an operand served by the reuse cache (same register in
the same slot as the previous FP64 instruction, flagged .reuse there) costs nothing.
   python tools/sass_reads.py <object or .so> <mangled kernel name> [window]"""
import re
import subprocess
import sys

obj, fun = sys.argv[1], sys.argv[2]
win = int(sys.argv[3]) if len(sys.argv) > 3 else 540
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = [m.group(1) for m in (re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", l) for l in txt.splitlines()) if m]
fp = lambda o: any(x in o for x in ("DFMA", "DADD", "DMUL"))
best = max(range(max(1, len(ins) - win)), key=lambda i: sum(1 for o in ins[i:i + win] if fp(o)))
loop = [o for o in ins[best:best + win] if fp(o)]
reads, n3, prev = 0, 0, None
for o in loop:
    body = o.split(" ", 1)[1] if o.startswith("@") else o
    name = body.split()[0]
    srcs = [x.strip() for x in body[len(name):].split(",")][1:]
    cur = [((m.group(1), ".reuse" in sx) if (m := re.search(r"R(\d+)", sx)) else None) for sx in srcs]
    cnt = sum(1 for k, c in enumerate(cur) if c and not (prev and k < len(prev) and prev[k] and prev[k][0] == c[0] and prev[k][1]))
    reads += cnt
    n3 += cnt >= 3
    prev = cur
cyc = sum(1 for _ in loop) * 2
print(f"{len(loop)} FP64 instructions, {reads / len(loop):.2f} register reads each, {n3} with three; "
      f"read-limited pipe rate {2 * len(loop) / max(2 * len(loop), reads):.2f} of peak (if a read costs a cycle)")
