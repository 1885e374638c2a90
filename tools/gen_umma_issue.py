#!/usr/bin/env python
"""Generates csrc/umma_issue.cuh: the single-asm-block `tcgen05.mma` issue sequences of the conv GEMM.

Why generated text: per-MMA issue cost decides the conv kernel.  An M128 x N112 x K16 MMA occupies the tensor pipe for
~60 cycles, and a C++ loop around single-MMA asm statements costs the issuing thread ~15 instructions (~130 cycles)
per MMA in uniform-register moves, predicates and branches (tools/micro/umma_bench.cu, stage_bench.cu).  All MMAs of
one pipeline stage -- (dy, k-chunk): `ntaps` dx taps x `ksteps` k-steps x `prod` products x `ntile` row tiles -- are
therefore issued from ONE straight-line asm block, one specialisation per stage shape, picked by a warp-uniform switch.

    prod = 3 : split operands x ~= hi + lo, products lo*hi + hi*lo + hi*hi   (fp32-grade)
    prod = 1 : hi*hi only                                                     (single pass)

Operands of every block (descriptor units are 16 bytes):
    %0 d_tmem0   accumulator of row tile 0          %1 bn        TMEM column stride tile -> tile
    %2 a_hi      smem descriptor (tile 0, tap 0, hi) %3 a_tile16  stride tile -> tile
    %4 a_box16   stride hi -> lo plane               %5 row16     stride tap -> tap (one smem row: the dx shift)
    %6 b_hi      descriptor (tap 0, hi plane)        %7 w_tap16   stride tap -> tap
    %8 w_plane16 stride hi -> lo plane               %9 idesc     %10 acc (0 = first MMA of each tile overwrites)

usage: python tools/gen_umma_issue.py <output.cuh>
"""
import sys

MAX_TILES = 2
TAPS = (1, 3)
KSTEPS = (1, 2, 3, 4)
ARGS = ("uint32_t d_tmem0, uint32_t bn, uint64_t a_hi, uint32_t a_tile16, uint32_t a_box16, uint32_t row16, "
        "uint64_t b_hi, uint32_t w_tap16, uint32_t w_plane16, uint32_t idesc, uint32_t acc")
CALL = "d_tmem0, bn, a_hi, a_tile16, a_box16, row16, b_hi, w_tap16, w_plane16, idesc, acc"
MMA = "tcgen05.mma.cta_group::1.kind::f16"


def block(prod, ntaps, ksteps, ntile):
    out = []
    emit = out.append
    tiles = range(ntile)
    emit(".reg .pred pacc, pt;")
    emit(".reg .b32 %s;" % ", ".join("d%d" % t for t in tiles))
    regs = ["bh", "bl", "at", "ab", "ar", "wt", "wb", "yh", "yl"]
    for t in tiles:
        regs += ["ah%d" % t, "al%d" % t, "xh%d" % t, "xl%d" % t]
    emit(".reg .b64 %s;" % ", ".join(regs))
    emit("setp.ne.b32 pacc, %10, 0;")
    emit("setp.eq.b32 pt, 0, 0;")
    emit("mov.b32 d0, %0;")
    for t in tiles[1:]:
        emit("add.u32 d%d, d%d, %%1;" % (t, t - 1))
    if prod == 3:
        emit("cvt.u64.u32 ab, %4;")
        emit("cvt.u64.u32 wb, %8;")
    if ntile > 1:
        emit("cvt.u64.u32 at, %3;")
    if ntaps > 1:
        emit("cvt.u64.u32 ar, %5;")
        emit("cvt.u64.u32 wt, %7;")
    emit("mov.b64 ah0, %2;")
    for t in tiles[1:]:
        emit("add.u64 ah%d, ah%d, at;" % (t, t - 1))
    emit("mov.b64 bh, %6;")
    if prod == 3:
        for t in tiles:
            emit("add.u64 al%d, ah%d, ab;" % (t, t))
        emit("add.u64 bl, bh, wb;")
    for s in range(ntaps):
        if s > 0:
            for t in tiles:
                emit("add.u64 ah%d, ah%d, ar;" % (t, t))
                if prod == 3:
                    emit("add.u64 al%d, al%d, ar;" % (t, t))
            emit("add.u64 bh, bh, wt;")
            if prod == 3:
                emit("add.u64 bl, bl, wt;")
        for k in range(ksteps):
            if k == 0:
                xh = ["ah%d" % t for t in tiles]
                xl = ["al%d" % t for t in tiles]
                yh, yl = "bh", "bl"
            else:
                off = 2 * k        # 16 halves = 32 bytes = 2 descriptor units per k-step
                for t in tiles:
                    emit("add.u64 xh%d, ah%d, %d;" % (t, t, off))
                    if prod == 3:
                        emit("add.u64 xl%d, al%d, %d;" % (t, t, off))
                emit("add.u64 yh, bh, %d;" % off)
                if prod == 3:
                    emit("add.u64 yl, bl, %d;" % off)
                xh = ["xh%d" % t for t in tiles]
                xl = ["xl%d" % t for t in tiles]
                yh, yl = "yh", "yl"
            first = "pacc" if (s == 0 and k == 0) else "pt"
            if prod == 3:
                for t in tiles:
                    emit("%s [d%d], %s, %s, %%9, %s;" % (MMA, t, xl[t], yh, first))
                for t in tiles:
                    emit("%s [d%d], %s, %s, %%9, pt;" % (MMA, t, xh[t], yl))
                for t in tiles:
                    emit("%s [d%d], %s, %s, %%9, pt;" % (MMA, t, xh[t], yh))
            else:
                for t in tiles:
                    emit("%s [d%d], %s, %s, %%9, %s;" % (MMA, t, xh[t], yh, first))
    return out


def function(prod, ntaps, ksteps, ntile):
    name = "umma_stage_x%d_t%dk%dn%d" % (prod, ntaps, ksteps, ntile)
    lines = ["__device__ __forceinline__ void %s(%s) {" % (name, ARGS), "    asm volatile(", '        "{\\n\\t"']
    for ins in block(prod, ntaps, ksteps, ntile):
        lines.append('        "%s\\n\\t"' % ins)
    lines.append('        "}"')
    lines.append('        ::"r"(d_tmem0), "r"(bn), "l"(a_hi), "r"(a_tile16), "r"(a_box16), "r"(row16), "l"(b_hi), '
                 '"r"(w_tap16), "r"(w_plane16),')
    lines.append('          "r"(idesc), "r"(acc)')
    lines.append('        : "memory");')
    lines.append("}")
    return name, "\n".join(lines)


def main(path):
    parts = ["// GENERATED by tools/gen_umma_issue.py -- do not edit (the generator's docstring explains the layout).",
             "#pragma once", "#include <stdint.h>", "", "namespace fsb {", "namespace umma {", "",
             "// one lane of a fully converged warp (always the same one)",
             "__device__ __forceinline__ bool elect_one() {", "    uint32_t ok;", "    asm volatile(",
             '        "{\\n\\t.reg .pred p;\\n\\t"', '        "elect.sync _|p, 0xffffffff;\\n\\t"',
             '        "selp.u32 %0, 1, 0, p;\\n\\t}"', '        : "=r"(ok));', "    return ok != 0;", "}", ""]
    for prod in (3, 1):
        cases = []
        for ntaps in TAPS:
            for ksteps in KSTEPS:
                for ntile in range(1, MAX_TILES + 1):
                    name, text = function(prod, ntaps, ksteps, ntile)
                    parts += [text, ""]
                    cases.append((ntaps * 100 + ksteps * 10 + ntile, name))
        parts.append("// all MMAs of one pipeline stage; returns false for a stage shape without a specialisation")
        parts.append("__device__ __forceinline__ bool umma_stage_x%d(%s, int ksteps, int ntile, int ntaps) {" % (prod, ARGS))
        parts.append("    switch (ntaps * 100 + ksteps * 10 + ntile) {")
        for code, name in cases:
            parts.append("        case %d: %s(%s); return true;" % (code, name, CALL))
        parts.append("        default: return false;")
        parts.append("    }")
        parts.append("}")
        parts.append("")
    parts += ["constexpr int kMaxTiles = %d;" % MAX_TILES, "", "}  // namespace umma", "}  // namespace fsb", ""]
    with open(path, "w") as f:
        f.write("\n".join(parts))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "umma_issue.cuh")
