"""GPU box: sweep the attention pipeline knobs with the native tool (tests/native/fmha_bench), check every variant's
whole-output checksum against pipeline 1 at a full, a ragged and a tiny length, and print the fastest correct
setting as shell exports. Usage: python profiles/r1_tools/fmha_pipe_sweep.py LOGFILE"""
import os
import re
import subprocess
import sys

TOOL = "tests/native/fmha_bench"
LENS = (11648, 11500, 300)


def run(L, pipe, tok, wa, poly, log):
    env = dict(os.environ, FX_FMHA_PIPE=str(pipe), FX_FMHA_TOKEN=str(tok), FX_FMHA_WARP_ARRIVE=str(wa),
               FX_FMHA_POLY=str(poly))
    try:
        out = subprocess.run([TOOL, str(L)], env=env, capture_output=True, text=True, timeout=120).stdout
    except subprocess.TimeoutExpired:
        out = "TIMEOUT\n"
    log.write(f"== fmha_bench L={L} pipe={pipe} token={tok} warp_arrive={wa} poly={poly}\n{out}")
    log.flush()
    r = {}
    m = re.search(r"launch 2: ([\d.]+) ms, ([\d.]+) TFLOP", out)
    if m:
        r["ms"], r["tflops"] = float(m[1]), float(m[2])
    m = re.search(r"rel-L2 ([\d.e+-]+) over .* checksum (\w+)", out)
    if m:
        r["rel"], r["sum"] = float(m[1]), m[2]
    return r


def main():
    log = open(sys.argv[1], "w")
    ref = {L: run(L, 1, 1, 0, 2, log) for L in LENS}
    best, best_ms = (1, 1, 0, 2), ref[LENS[0]].get("ms", 1e9)
    table = [("1 1 0 2", ref[LENS[0]])]
    for pipe, tok, wa in ((1, 0, 0), (2, 1, 0), (2, 0, 0), (3, 1, 0), (3, 0, 0), (2, 1, 1), (3, 1, 1), (3, 0, 1)):
        rs = {L: run(L, pipe, tok, wa, 2, log) for L in LENS}
        ok = all(rs[L].get("sum") is not None and rs[L].get("sum") == ref[L].get("sum") and rs[L]["rel"] < 8e-3
                 for L in LENS)
        table.append((f"{pipe} {tok} {wa} 2 {'ok' if ok else 'MISMATCH'}", rs[LENS[0]]))
        if ok and rs[LENS[0]]["ms"] < best_ms:
            best, best_ms = (pipe, tok, wa, 2), rs[LENS[0]]["ms"]
    # exp2 split on the winning pipeline (different arithmetic: accuracy gate only)
    for poly in (0, 3, 4):
        rs = {L: run(L, best[0], best[1], best[2], poly, log) for L in (LENS[0], LENS[2])}
        ok = all("rel" in rs[L] and rs[L]["rel"] < 8e-3 for L in rs)
        table.append((f"{best[0]} {best[1]} {best[2]} {poly} {'ok' if ok else 'BAD'}", rs[LENS[0]]))
        if ok and rs[LENS[0]]["ms"] < best_ms:
            best, best_ms = (best[0], best[1], best[2], poly), rs[LENS[0]]["ms"]
    for name, r in table:
        log.write(f"# pipe tok wa poly = {name}: {r}\n")
        print(f"# pipe tok wa poly = {name}: {r}", file=sys.stderr)
    log.write(f"# chosen {best} {best_ms} ms\n")
    print(f"export FX_FMHA_PIPE={best[0]} FX_FMHA_TOKEN={best[1]} FX_FMHA_WARP_ARRIVE={best[2]} FX_FMHA_POLY={best[3]}")


if __name__ == "__main__":
    main()
