#!/usr/bin/env python
"""Generates housescan_b200/csrc/k_eval_group_gen.cuh: the PTX blocks that evaluate one 48-byte group (4 points) of the
cuboid-sums kernel.  Generated rather than hand-numbered because each block has ~45 operands.

    python tools/gen_eval_asm.py          # rewrites the header in place (the header is committed)

Variants (all bit-identical in their results; they differ in instruction mix):
  add_group_v1  packed products (FMUL2 on the AoS register pairs as they come out of LDS.128) and packed +/- wall offsets
                (FADD2); scalar sums; predicated scalar accumulation.
  add_group_v2  v1 + packed predicated accumulation (FFMA2/FADD2 on accumulator pairs).
"""
import os

ACC = ["f", "T0", "T1", "T2", "M0", "M1", "M2", "B00", "B01", "B02", "B10", "B11", "B12", "B20", "B21", "B22", "C1", "C2", "Cm0", "Cm1", "Cm2"]
ACC_EXPR = {"f": "c.f", "C1": "c.C1", "C2": "c.C2"}
for j in range(3):
    ACC_EXPR[f"T{j}"] = f"c.T[{j}]"
    ACC_EXPR[f"M{j}"] = f"c.M[{j}]"
    ACC_EXPR[f"Cm{j}"] = f"c.Cm[{j}]"
    for q in range(3):
        ACC_EXPR[f"B{j}{q}"] = f"c.B[{j}][{q}]"


class Ops:
    def __init__(self):
        self.outs, self.ins, self.idx = [], [], {}

    def out(self, name, expr, kind="f"):
        self.idx[name] = len(self.outs)
        self.outs.append((kind, expr))

    def inp(self, name, expr, kind="f"):
        self.idx[name] = ("in", len(self.ins))
        self.ins.append((kind, expr))

    def ref(self, name):
        v = self.idx[name]
        return f"%{v}" if isinstance(v, int) else f"%{len(self.outs) + v[1]}"


def dots_and_offsets(o, L):
    """t_j of the 4 points of the group and the signed distances to both walls of every axis (spJ_k, smJ_k)."""
    for j in range(3):
        kab, kca, kbc = o.ref(f"Kab{j}"), o.ref(f"Kca{j}"), o.ref(f"Kbc{j}")
        L.append(f"mul.rn.f32x2 q0, {o.ref('w0')}, {kab};\n mul.rn.f32x2 q1, {o.ref('w1')}, {kca};\n mul.rn.f32x2 q2, {o.ref('w2')}, {kbc};")
        L.append(f"mul.rn.f32x2 q3, {o.ref('w3')}, {kab};\n mul.rn.f32x2 q4, {o.ref('w4')}, {kca};\n mul.rn.f32x2 q5, {o.ref('w5')}, {kbc};")
        # point 0: (q0.lo + q0.hi) + q1.lo ; point 1: (q1.hi + q2.lo) + q2.hi ; same for 2, 3 with q3..q5
        L.append("mov.b64 {l0, h0}, q0;\n mov.b64 {l1, h1}, q1;\n mov.b64 {l2, h2}, q2;")
        L.append("add.rn.f32 u, l0, h0;\n add.rn.f32 t0, u, l1;\n add.rn.f32 u, h1, l2;\n add.rn.f32 t1, u, h2;")
        L.append("mov.b64 {l0, h0}, q3;\n mov.b64 {l1, h1}, q4;\n mov.b64 {l2, h2}, q5;")
        L.append("add.rn.f32 u, l0, h0;\n add.rn.f32 t2, u, l1;\n add.rn.f32 u, h1, l2;\n add.rn.f32 t3, u, h2;")
        L.append("mov.b64 q0, {t0, t1};\n mov.b64 q1, {t2, t3};")
        L.append(f"add.rn.f32x2 q2, q0, {o.ref(f'NDP{j}')};\n add.rn.f32x2 q3, q1, {o.ref(f'NDP{j}')};")
        L.append(f"add.rn.f32x2 q4, q0, {o.ref(f'DM{j}')};\n add.rn.f32x2 q5, q1, {o.ref(f'DM{j}')};")
        L.append(f"mov.b64 {{sp{j}_0, sp{j}_1}}, q2;\n mov.b64 {{sp{j}_2, sp{j}_3}}, q3;\n mov.b64 {{sm{j}_0, sm{j}_1}}, q4;\n mov.b64 {{sm{j}_2, sm{j}_3}}, q5;")


def select_point(k, L, dup=False):
    """sides and nearest axis of point k -> s0..s2 (and sd0..sd2 duplicates), p0..p2, predicates E0..E2"""
    for j in range(3):
        L.append(f"abs.f32 asp, sp{j}_{k};\n abs.f32 asm_, sm{j}_{k};\n setp.lt.f32 P, asm_, asp;")
        L.append(f"selp.f32 s{j}, sm{j}_{k}, sp{j}_{k}, P;\n selp.f32 p{j}, 0f3F800000, 0f00000000, P;")
        if dup:
            L.append(f"selp.f32 sd{j}, sm{j}_{k}, sp{j}_{k}, P;")
    L.append("abs.f32 a0, s0;\n abs.f32 a1, s1;\n abs.f32 a2, s2;")
    L.append("setp.lt.f32 Q1, a1, a0;\n selp.f32 a01, a1, a0, Q1;\n setp.lt.f32 E2, a2, a01;")
    L.append("setp.lt.and.f32 E1, a1, a0, !E2;\n setp.geu.and.f32 E0, a1, a0, !E2;")


def accumulate_scalar(o, k, L):
    xyz = [f"x{k}", f"y{k}", f"z{k}"]
    for j in range(3):
        E = f"@E{j}"
        L.append(f"{E} fma.rn.f32 {o.ref('f')}, s{j}, s{j}, {o.ref('f')};\n {E} add.rn.f32 {o.ref(f'T{j}')}, {o.ref(f'T{j}')}, s{j};")
        L.append(f"{E} fma.rn.f32 {o.ref(f'M{j}')}, s{j}, p{j}, {o.ref(f'M{j}')};\n {E} add.rn.f32 {o.ref(f'Cm{j}')}, {o.ref(f'Cm{j}')}, p{j};")
        for q in range(3):
            L.append(f"{E} fma.rn.f32 {o.ref(f'B{j}{q}')}, s{j}, {xyz[q]}, {o.ref(f'B{j}{q}')};")
        if j:
            L.append(f"{E} add.rn.f32 {o.ref(f'C{j}')}, {o.ref(f'C{j}')}, 0f3F800000;")


def emit(name, o, L, decl):
    body = "\n".join(L)
    lines = ['      "{\\n"'] + [f'      "{decl}\\n"']
    for ln in body.split("\n"):
        lines.append(f'      "{ln.strip()}\\n"')
    lines.append('      "}\\n"')
    outs = ", ".join(f'"+{k}"({e})' for k, e in o.outs)
    ins = ", ".join(f'"{k}"({e})' for k, e in o.ins)
    return "\n".join(lines) + f"\n      : {outs}\n      : {ins});"


def gen_v1(name="add_group_v1", mode="full"):
    o = Ops()
    for a in ACC:
        o.out(a, ACC_EXPR[a])
    for i in range(6):
        o.inp(f"w{i}", f"w{i}", "l")
    for j in range(3):
        for nm in ("Kab", "Kca", "Kbc"):
            o.inp(f"{nm}{j}", f"R.{nm}[{j}]", "l")
    for j in range(3):
        o.inp(f"NDP{j}", f"R.ndp[{j}]", "l")
    for j in range(3):
        o.inp(f"DM{j}", f"R.dm[{j}]", "l")
    L = []
    L.append(f"mov.b64 {{x0, y0}}, {o.ref('w0')};\n mov.b64 {{z0, x1}}, {o.ref('w1')};\n mov.b64 {{y1, z1}}, {o.ref('w2')};")
    L.append(f"mov.b64 {{x2, y2}}, {o.ref('w3')};\n mov.b64 {{z2, x3}}, {o.ref('w4')};\n mov.b64 {{y3, z3}}, {o.ref('w5')};")
    dots_and_offsets(o, L)
    for k in range(4):
        select_point(k, L)
        if mode == "full":
            accumulate_scalar(o, k, L)
        elif mode == "nopred":
            L2 = []
            accumulate_scalar(o, k, L2)
            L += [x.replace("@E0 ", "").replace("@E1 ", "").replace("@E2 ", "") for x in L2]
        elif mode == "front":
            for j in range(3):
                L.append(f"@E{j} add.rn.f32 {o.ref('f')}, {o.ref('f')}, s{j};\n add.rn.f32 {o.ref(f'T{j}')}, {o.ref(f'T{j}')}, p{j};")
    decl = (".reg .pred P, Q1, E0, E1, E2;\\n"
            ".reg .f32 x0, y0, z0, x1, y1, z1, x2, y2, z2, x3, y3, z3, l0, h0, l1, h1, l2, h2, u, t0, t1, t2, t3;\\n"
            ".reg .f32 asp, asm_, s0, s1, s2, p0, p1, p2, a0, a1, a2, a01;\\n"
            ".reg .b64 q0, q1, q2, q3, q4, q5;\\n"
            ".reg .f32 " + ", ".join(f"sp{j}_{k}, sm{j}_{k}" for j in range(3) for k in range(4)) + ";")
    return ("// packed products / wall offsets, scalar sums, predicated scalar accumulation\n"
            "__device__ __forceinline__ void " + name + "(ChainsP& c, const RoomK2& R, unsigned long long w0, unsigned long long w1,\n"
            "    unsigned long long w2, unsigned long long w3, unsigned long long w4, unsigned long long w5) {\n  asm(\n" + emit("v1", o, L, decl) + "\n}\n")


HEADER = '''// GENERATED by tools/gen_eval_asm.py — do not edit by hand.
// PTX blocks evaluating one 48-byte group (4 points) of the cuboid-sums kernel; semantics in k_eval_point.cuh.
#pragma once

namespace hsk {

// one room's constants as packed register pairs: for axis j with + normal (a, b, c):
//   Kab = (a, b), Kca = (c, a), Kbc = (b, c) multiply the AoS pairs (x0,y0) (z0,x1) (y1,z1) | (x2,y2) (z2,x3) (y3,z3);
//   ndp = (-d+, -d+), dm = (d-, d-)
struct RoomK2 {
  unsigned long long Kab[3], Kca[3], Kbc[3], ndp[3], dm[3];
};
__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ RoomK2 make_room_k2(const RoomK& r) {
  RoomK2 o;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    o.Kab[j] = pack_f2(r.n[j][0], r.n[j][1]);
    o.Kca[j] = pack_f2(r.n[j][2], r.n[j][0]);
    o.Kbc[j] = pack_f2(r.n[j][1], r.n[j][2]);
    o.ndp[j] = pack_f2(-r.dp[j], -r.dp[j]);
    o.dm[j] = pack_f2(r.dm[j], r.dm[j]);
  }
  return o;
}

'''

if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "housescan_b200", "csrc", "k_eval_group_gen.cuh")
    with open(out, "w") as fh:
        fh.write(HEADER + gen_v1() + "\n}  // namespace hsk\n")
    print("wrote", out)
    abl = os.path.join(root, "tools", "ubench4_ablations.cuh")
    with open(abl, "w") as fh:
        fh.write("// GENERATED by tools/gen_eval_asm.py: timing-only ablations of add_group_v1 (results are NOT the product's)\n#pragma once\nnamespace hsk {\n"
                 + gen_v1("add_group_nopred", "nopred") + gen_v1("add_group_front", "front") + "}\n")
    print("wrote", abl)
