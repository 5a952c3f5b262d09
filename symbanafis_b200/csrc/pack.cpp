// pack.cpp -- device program (indices, run-time kinds) -> micro-instruction image (byte offsets,
// kind-specialised micro-ops).  Pure re-encoding: the same operations run on the same values in the
// same order; commutative operands may swap places (IEEE add/mul commute bit for bit).  A peephole then fuses
// chains of micro-ops whose intermediate results only travel through the accumulator (uop.h U_FM .. U_FT).
#include "pack.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <sstream>

#include "../../include/saf_b200.h"

namespace saf {

namespace {

struct Dec { // one decoded device instruction
    uint32_t op, prefix, fd, fa, fb, fc, imm;
};

struct UI {
    uint32_t uop = U_END;
    int32_t d = NO_DEST;
    uint32_t a = 0, b = 0, c = 0, imm = 0, kinds = 0;
};

} // namespace

int pack_program(const DeviceProgram &dp, int P, PackedProgram &out, std::string &err, int block_threads) {
    out = PackedProgram();
    out.points_per_thread = P;
    out.block_threads = block_threads;
    out.layout_id = block_threads == BLOCK_THREADS ? P : 1000 * P + block_threads;
    if (block_threads != BLOCK_THREADS && ((P != 2 && P != 4) || block_threads % 32 != 0 || block_threads > WIDE_MAX_THREADS)) {
        err = "internal: wide blocks exist for two or four points per thread only";
        return SAF_ERR_INVALID_ARGUMENT;
    }
    const uint32_t SB = slot_bytes(P, block_threads);

    // ---- decode ----
    std::vector<Dec> code;
    bool needs_staging = false;
    for (size_t i = 0; i < dp.code.size();) {
        const uint64_t lo = dp.code[i++];
        const uint32_t x = (uint32_t)(lo & 0xffu);
        uint32_t op, ka, kb;
        decode_xop(x, &op, &ka, &kb);
        Dec d{op, (uint32_t)((lo >> 8) & 0xffu), (uint32_t)((lo >> 16) & 0xffffu), (uint32_t)((lo >> 32) & 0xffffu),
              (uint32_t)(lo >> 48), FIELD_NONE, 0};
        if (xop_has_hi(x)) {
            if (i >= dp.code.size()) {
                err = "internal: truncated device program";
                return SAF_ERR_INVALID_PROGRAM;
            }
            const uint64_t hi = dp.code[i++];
            d.fc = (uint32_t)(hi & 0xffffu);
            d.imm = (uint32_t)(hi >> 32);
        }
        if (op == D_BUILTIN3 || op == D_BUILTIN4 || op == D_ARG2) needs_staging = true;
        code.push_back(d);
        if (op == D_END) break;
    }

    // ---- register file and constant table ----
    uint32_t n_slots = dp.n_slots;
    out.acc_off = (int32_t)(n_slots * SB);
    n_slots += 1;
    if (needs_staging) {
        out.x0_off = (int32_t)(n_slots * SB);
        out.x1_off = (int32_t)((n_slots + 1) * SB);
        n_slots += 2;
    }
    out.n_slots = n_slots;
    if ((uint64_t)n_slots * SB >= 0x7fffffffull) {
        err = "register file offset overflow";
        return SAF_ERR_UNSUPPORTED;
    }
    std::vector<double> K(dp.consts.begin(), dp.consts.end());
    std::map<uint64_t, uint32_t> mul_pair; // bits(k1) -> byte offset of the pair {k1, -0.0}
    // slot renaming of the current pass (identity for copy A, parameter <-> result swap for copy B)
    std::vector<uint32_t> slot_map(dp.n_slots);
    for (uint32_t i = 0; i < dp.n_slots; ++i) slot_map[i] = i;
    auto soff = [&](uint32_t f) {
        const uint32_t i = field_index(f);
        return (i < slot_map.size() ? slot_map[i] : i) * SB;
    };
    auto koff = [&](uint32_t f) { return field_index(f) * 8u; };
    auto off_of = [&](uint32_t f) -> uint32_t {
        switch (field_kind(f)) {
        case KIND_S: return soff(f);
        case KIND_K: return koff(f);
        default: return 0;
        }
    };
    auto dest_of = [&](uint32_t f) -> int32_t { return field_kind(f) == KIND_S ? (int32_t)soff(f) : NO_DEST; };

    auto kinds3 = [](uint32_t fa, uint32_t fb, uint32_t fc) {
        return field_kind(fa) | (field_kind(fb) << 2) | (field_kind(fc) << 4);
    };

    // converts the decoded program under the current slot_map; returns 0 or an error code
    auto build = [&](std::vector<UI> &U) -> int {
    for (const Dec &d : code) {
        UI u;
        u.d = dest_of(d.fd);
        const uint32_t ka = field_kind(d.fa), kb = field_kind(d.fb), kc = field_kind(d.fc);
        const uint32_t op = d.op;
        if (op == D_END) {
            u.uop = U_END;
        } else if (op == D_LOADK) {
            u.uop = U_LOADK;
            u.a = koff(d.fa);
        } else if (op >= D_FIRST_LIGHT_UNARY && op < D_FIRST_LIGHT_UNARY + D_N_LIGHT_UNARY) {
            if (ka == KIND_K) { // not produced by the lowering; keep the packer total anyway
                UI l;
                l.uop = U_LOADK;
                l.a = koff(d.fa);
                U.push_back(l);
            }
            u.uop = U_U1 + 2 * (op - D_FIRST_LIGHT_UNARY) + (ka == KIND_S ? 0u : 1u);
            u.a = off_of(d.fa);
        } else if (op == D_SIN || op == D_COS || op == D_EXP || op == D_LN || op == D_RECIPEXPM1 || op == D_EXPSQR ||
                   op == D_EXPSQRNEG || op == D_EXPNEG || op == D_SINCOS) {
            uint32_t pre_off = NO_PREFIX;
            if (d.prefix == PREFIX_FMA) {
                pre_off = koff(d.fb); // K[b], K[b+1] are adjacent in the device constant table
            } else if (d.prefix == PREFIX_MUL) {
                const double k1 = dp.consts[field_index(d.fb)];
                uint64_t bits;
                std::memcpy(&bits, &k1, 8);
                auto it = mul_pair.find(bits);
                if (it == mul_pair.end()) {
                    const uint32_t o = (uint32_t)K.size() * 8u;
                    K.push_back(k1);
                    K.push_back(-0.0); // fma(x, k1, -0.0) == x * k1 for every x, zeros and NaN included
                    it = mul_pair.emplace(bits, o).first;
                }
                pre_off = it->second;
            }
            if (op == D_RECIPEXPM1) { // cold
                u.uop = U_G + G_RECIPEXPM1;
                u.a = off_of(d.fa);
                u.b = pre_off;
                u.kinds = kinds3(d.fa, FIELD_NONE, FIELD_NONE);
            } else {
                static const uint32_t H[] = {H1_SIN, H1_COS, H1_EXP, H1_LN, 0, H1_EXPSQR, H1_EXPSQRNEG, H1_EXPNEG, H1_SINCOS};
                const uint32_t h = H[op - D_SIN];
                // the hot heavy handlers read their operand from a slot only
                if (ka == KIND_K) { // constant argument: materialise it through the reserved slot
                    UI l;
                    l.uop = U_LOADK;
                    l.a = koff(d.fa);
                    l.d = out.acc_off;
                    U.push_back(l);
                    u.a = (uint32_t)out.acc_off;
                } else if (ka == KIND_ACC) { // the producer is the previous micro-op: make it store its result
                    if (U.empty()) {
                        err = "internal: accumulator operand without a producer";
                        return SAF_ERR_INVALID_PROGRAM;
                    }
                    if (U.back().d == NO_DEST) U.back().d = out.acc_off;
                    u.a = (uint32_t)U.back().d;
                } else {
                    u.a = off_of(d.fa);
                }
                u.uop = U_H1 + 2 * h + (pre_off != NO_PREFIX ? 1u : 0u);
                u.b = pre_off;
                if (op == D_SINCOS) u.c = (uint32_t)dest_of(d.fc);
            }
        } else if (op == D_ADD || op == D_MUL || op == D_NEGMUL) {
            const uint32_t i = op == D_ADD ? BC_ADD : op == D_MUL ? BC_MUL : BC_NEGMUL;
            uint32_t fa = d.fa, fb = d.fb;
            auto rank = [](uint32_t k) { return k == KIND_S ? 0 : k == KIND_K ? 1 : 2; };
            if (rank(field_kind(fa)) > rank(field_kind(fb))) std::swap(fa, fb);
            const uint32_t a_ = field_kind(fa), b_ = field_kind(fb);
            uint32_t v;
            if (a_ == KIND_S && b_ == KIND_S) v = BC_SS;
            else if (a_ == KIND_S && b_ == KIND_K) v = BC_SK;
            else if (a_ == KIND_S && b_ == KIND_ACC) v = BC_SA;
            else if (a_ == KIND_K && b_ == KIND_ACC) v = BC_KA;
            else if (a_ == KIND_ACC && b_ == KIND_ACC) v = BC_AA;
            else {
                err = "internal: constant-constant light binary operation in the device program";
                return SAF_ERR_INVALID_PROGRAM;
            }
            u.uop = U_BC + 5 * i + v;
            u.a = off_of(fa);
            u.b = off_of(fb);
        } else if (op == D_SUB || op == D_DIV) {
            const uint32_t i = op == D_SUB ? BN_SUB : BN_DIV;
            int v = -1;
            if (ka == KIND_S && kb == KIND_S) v = BN_SS;
            else if (ka == KIND_S && kb == KIND_K) v = BN_SK;
            else if (ka == KIND_K && kb == KIND_S) v = BN_KS;
            else if (ka == KIND_S && kb == KIND_ACC) v = BN_SA;
            else if (ka == KIND_ACC && kb == KIND_S) v = BN_AS;
            else if (ka == KIND_K && kb == KIND_ACC) v = BN_KA;
            else if (ka == KIND_ACC && kb == KIND_K) v = BN_AK;
            else if (ka == KIND_ACC && kb == KIND_ACC) v = BN_AA;
            if (v >= 0) {
                u.uop = U_BN + 8 * i + (uint32_t)v;
            } else if (op == D_DIV) {
                u.uop = U_G + G_DIV;
                u.kinds = kinds3(d.fa, d.fb, FIELD_NONE);
            } else {
                err = "internal: constant-constant subtraction in the device program";
                return SAF_ERR_INVALID_PROGRAM;
            }
            u.a = off_of(d.fa);
            u.b = off_of(d.fb);
        } else if (op >= D_FIRST_TERNARY && op < D_FIRST_TERNARY + D_N_TERNARY) {
            const uint32_t t = op - D_FIRST_TERNARY;
            uint32_t fa = d.fa, fb = d.fb;
            auto rank = [](uint32_t k) { return k == KIND_S ? 0 : k == KIND_K ? 1 : 2; };
            if (rank(field_kind(fa)) > rank(field_kind(fb))) std::swap(fa, fb); // a * b commutes
            const uint32_t a_ = field_kind(fa), b_ = field_kind(fb);
            int v = -1;
            if (a_ == KIND_S && b_ == KIND_S) v = kc == KIND_S ? (int)T_SSS : kc == KIND_K ? (int)T_SSK : (int)T_SSA;
            else if (a_ == KIND_S && b_ == KIND_K) v = kc == KIND_S ? (int)T_SKS : kc == KIND_K ? (int)T_SKK : (int)T_SKA;
            else if (a_ == KIND_S && b_ == KIND_ACC && kc != KIND_ACC) v = kc == KIND_S ? (int)T_SAS : (int)T_SAK;
            else if (a_ == KIND_K && b_ == KIND_ACC && kc != KIND_ACC) v = kc == KIND_S ? (int)T_KAS : (int)T_KAK;
            if (v >= 0) {
                u.uop = U_T + 10 * t + (uint32_t)v;
                u.a = off_of(fa);
                u.b = off_of(fb);
                u.c = off_of(d.fc);
            } else {
                u.uop = U_G + G_FMA + t;
                u.a = off_of(d.fa);
                u.b = off_of(d.fb);
                u.c = off_of(d.fc);
                u.kinds = kinds3(d.fa, d.fb, d.fc);
            }
        } else {
            // generic micro-ops: kinds decoded at run time
            u.a = off_of(d.fa);
            u.b = off_of(d.fb);
            u.imm = d.imm;
            u.kinds = kinds3(d.fa, d.fb, FIELD_NONE);
            switch (op) {
            case D_POW: u.uop = U_G + G_POW; break;
            case D_POWI: u.uop = U_G + G_POWI; break;
            case D_BUILTIN1: u.uop = U_G + G_BUILTIN1; break;
            case D_BUILTIN2: u.uop = U_G + G_BUILTIN2; break;
            case D_BUILTIN3: u.uop = U_G + G_BUILTIN3; break;
            case D_BUILTIN4: u.uop = U_G + G_BUILTIN4; break;
            case D_ARG2: u.uop = U_G + G_ARG2; u.d = NO_DEST; break;
            default: {
                std::ostringstream os;
                os << "internal: device op " << op << " has no micro-op";
                err = os.str();
                return SAF_ERR_INVALID_PROGRAM;
            }
            }
        }
        U.push_back(u);
    }
    return SAF_OK;
    }; // build

    // ---- fused chains (uop.h U_FM / U_FQ): multiply | square, then "K * ACC" and / or "S + ACC" ----
    static const bool fuse_chains = []() {
        const char *v = getenv("SAF_FUSE"); // tuning knob: SAF_FUSE=0 keeps one micro-op per device instruction
        return !(v && v[0] == '0');
    }();
    auto fuse = [&](std::vector<UI> &U) {
        if (!fuse_chains) return;
        const uint32_t MULKA = U_BC + 5 * BC_MUL + BC_KA, ADDSA = U_BC + 5 * BC_ADD + BC_SA;
        const uint32_t SQ = U_U1 + 2 * (D_SQUARE - D_FIRST_LIGHT_UNARY);
        std::vector<UI> F;
        F.reserve(U.size());
        for (size_t i = 0; i < U.size();) {
            const UI &p = U[i];
            const bool is_mul = p.uop >= U_BC + 5 * BC_MUL && p.uop < U_BC + 5 * BC_MUL + 5;
            const bool is_sq = p.uop == SQ || p.uop == SQ + 1;
            // S = trig1(pre), then trig2(pre) combined with that S by an add / fma that overwrites it (uop.h U_FD)
            if (P == 4 && i + 2 < U.size() && p.d != NO_DEST && (p.uop == U_H1 + 2 * H1_SIN + 1 || p.uop == U_H1 + 2 * H1_COS + 1)) {
                const UI &q = U[i + 1], &r = U[i + 2];
                const uint32_t FMASKA = U_T + 10 * T_FMA + T_SKA;
                if ((q.uop == U_H1 + 2 * H1_SIN + 1 || q.uop == U_H1 + 2 * H1_COS + 1) && q.d == NO_DEST &&
                    (r.uop == ADDSA || r.uop == FMASKA) && r.a == (uint32_t)p.d && r.d == p.d && q.a != (uint32_t)p.d) {
                    UI f = p;
                    const uint32_t t1 = (p.uop - U_H1) / 2 == H1_SIN ? 0u : 1u, t2 = (q.uop - U_H1) / 2 == H1_SIN ? 0u : 1u;
                    f.uop = U_FD + 4 * t1 + 2 * t2 + (r.uop == FMASKA ? 1u : 0u);
                    f.imm = q.a;
                    f.kinds = q.b;
                    if (r.uop == FMASKA) f.c = r.b;
                    f.d = r.d;
                    F.push_back(f);
                    i += 3;
                    continue;
                }
            }
            // sin / cos then S + ACC or fma(S, K, ACC): four-points-per-thread layout only (uop.h U_FT)
            if (P == 4 && p.d == NO_DEST && i + 1 < U.size() && p.uop >= U_H1 + 2 * H1_SIN && p.uop < U_H1 + 2 * H1_COS + 2) {
                const UI &q = U[i + 1];
                const uint32_t trig = (p.uop - U_H1) / 2 == H1_SIN ? 0u : 1u, prefixed = (p.uop - U_H1) & 1u;
                const uint32_t FMASKA = U_T + 10 * T_FMA + T_SKA;
                if (q.uop == ADDSA || q.uop == FMASKA) {
                    UI f = p;
                    f.uop = U_FT + 4 * trig + 2 * (q.uop == FMASKA ? FT_FMA : FT_ADDS) + prefixed;
                    f.imm = q.a;
                    if (q.uop == FMASKA) f.c = q.b;
                    f.d = q.d;
                    F.push_back(f);
                    i += 2;
                    continue;
                }
            }
            // t = S + K (kept or not), then t * t [+ S] or K * (K * t)
            if (p.uop == U_BC + 5 * BC_ADD + BC_SK && i + 1 < U.size()) {
                const UI &q = U[i + 1];
                UI f = p;
                f.kinds = (uint32_t)p.d; // where t itself is kept
                size_t used = 0;
                if (q.uop == SQ + 1) {
                    if (q.d == NO_DEST && i + 2 < U.size() && U[i + 2].uop == ADDSA) {
                        f.uop = U_FS + FS_SQ_ADDS;
                        f.imm = U[i + 2].a;
                        f.d = U[i + 2].d;
                        used = 3;
                    } else {
                        f.uop = U_FS + FS_SQ;
                        f.d = q.d;
                        used = 2;
                    }
                } else if (q.uop == MULKA && q.d == NO_DEST && i + 2 < U.size() && U[i + 2].uop == MULKA) {
                    f.uop = U_FS + FS_MULK2;
                    f.c = q.a;
                    f.imm = U[i + 2].a;
                    f.d = U[i + 2].d;
                    used = 3;
                }
                if (used) {
                    F.push_back(f);
                    i += used;
                    continue;
                }
            }
            if ((is_mul || is_sq) && p.d == NO_DEST && i + 1 < U.size()) {
                const UI &q = U[i + 1];
                UI f = p;
                size_t used = 0;
                if (q.uop == MULKA) {
                    if (is_mul && q.d == NO_DEST && i + 2 < U.size() && U[i + 2].uop == ADDSA) {
                        f.uop = U_FM + 3 * (p.uop - (U_BC + 5 * BC_MUL)) + F_MULK_ADDS;
                        f.c = q.a;
                        f.imm = U[i + 2].a;
                        f.d = U[i + 2].d;
                        used = 3;
                    } else {
                        f.uop = is_mul ? U_FM + 3 * (p.uop - (U_BC + 5 * BC_MUL)) + F_MULK : U_FQ + 2 * (p.uop - SQ) + F_MULK;
                        f.c = q.a;
                        f.d = q.d;
                        used = 2;
                    }
                } else if (q.uop == ADDSA) {
                    f.uop = is_mul ? U_FM + 3 * (p.uop - (U_BC + 5 * BC_MUL)) + F_ADDS : U_FQ + 2 * (p.uop - SQ) + F_ADDS;
                    f.imm = q.a;
                    f.d = q.d;
                    used = 2;
                }
                if (used) {
                    F.push_back(f);
                    i += used;
                    continue;
                }
            }
            F.push_back(p);
            ++i;
        }
        // two fused multiply chains back to back, accumulated into their own destination (uop.h U_F5)
        std::vector<UI> G;
        G.reserve(F.size());
        const uint32_t FM_SK_MULK = U_FM + 3 * BC_SK + F_MULK, FM_SA_MULK_ADDS = U_FM + 3 * BC_SA + F_MULK_ADDS;
        for (size_t i = 0; i < F.size(); ++i) {
            const UI &p = F[i];
            if (p.uop == FM_SK_MULK && p.d == NO_DEST && i + 1 < F.size() && F[i + 1].uop == FM_SA_MULK_ADDS &&
                F[i + 1].d >= 0 && (uint32_t)F[i + 1].d == F[i + 1].imm) {
                const UI &q = F[i + 1];
                UI f = p;           // a = S1, b = K1, c = K2
                f.uop = U_F5;
                f.imm = q.a;        // S2
                f.kinds = q.c;      // K3
                f.d = q.d;          // S3 = destination
                G.push_back(f);
                ++i;
                continue;
            }
            G.push_back(p);
        }
        U.swap(G);
    };

    std::vector<UI> U;
    int rc = build(U);
    if (rc != SAF_OK) return rc;
    fuse(U);

    // ---- double buffering for iterated maps --------------------------------------------------------------
    // Between two iterations of saf_iterate_* parameter j must receive result j.  Instead of copying slots,
    // the image carries a second copy of the program in which parameter slot and result slot of every j have
    // swapped names: copy A leaves its results where copy B expects its parameters and vice versa, and the
    // U_END of each copy continues with the other one.  Possible when the pairing is a bijection on slots.
    const size_t n_res = dp.result_field.size();
    bool dbuf = n_res == dp.param_count && n_res > 0;
    if (dbuf) {
        std::vector<int> seen(dp.n_slots, 0);
        for (uint32_t j = 0; j < dp.param_count && dbuf; ++j) {
            const uint32_t f = dp.result_field[j];
            if (field_kind(f) != KIND_S) dbuf = false;                  // constant result: needs the copy loop
            else if (dp.param_slot[j] == NO_SLOT) continue;             // nothing to feed back
            else if (seen[field_index(f)]++ || seen[dp.param_slot[j]]++) dbuf = false;
        }
        // a result slot that is fed back must not also be the result of a parameter that is not
        for (uint32_t j = 0; j < dp.param_count && dbuf; ++j)
            if (dp.param_slot[j] == NO_SLOT && seen[field_index(dp.result_field[j])]) dbuf = false;
    }
    const size_t n_a = U.size(); // micro-instructions of copy A incl. its U_END
    if (dbuf) {
        for (uint32_t j = 0; j < dp.param_count; ++j) {
            if (dp.param_slot[j] == NO_SLOT) continue;
            const uint32_t r = field_index(dp.result_field[j]), q = dp.param_slot[j];
            slot_map[q] = r;
            slot_map[r] = q;
        }
        std::vector<UI> B;
        rc = build(B);
        if (rc != SAF_OK) return rc;
        fuse(B);
        U.insert(U.end(), B.begin(), B.end());
    }

    // ---- windows: the last slot of every code window is U_NEXTWIN (uop.h) ----
    // index_of[i] = slot of logical instruction i in the windowed stream
    std::vector<UI> Wd;
    std::vector<size_t> index_of(U.size());
    for (size_t i = 0; i < U.size(); ++i) {
        if (Wd.size() % CODE_WINDOW_INSTR == CODE_WINDOW_INSTR - 1) {
            UI nw;
            nw.uop = U_NEXTWIN;
            Wd.push_back(nw);
        }
        index_of[i] = Wd.size();
        Wd.push_back(U[i]);
    }
    const size_t start_a = index_of[0], start_b = dbuf ? index_of[n_a] : index_of[0];
    const size_t end_a = index_of[n_a - 1], end_b = index_of[U.size() - 1];
    // U_END.a = code offset where the next iteration starts
    Wd[end_a].a = (uint32_t)(start_b * UINSTR_BYTES);
    if (dbuf) Wd[end_b].a = (uint32_t)(start_a * UINSTR_BYTES);

    // ---- image: [constants, padded to 32 bytes][code] ----
    const size_t k_words = (K.size() + 3) & ~(size_t)3;
    out.image.assign(k_words + Wd.size() * UINSTR_WORDS, 0);
    for (size_t i = 0; i < K.size(); ++i) std::memcpy(&out.image[i], &K[i], 8);
    // word 0 carries the opcode of the NEXT micro-instruction to execute (uop.h): the dispatch of instruction
    // i + 1 then does not wait for the fetch of instruction i + 1.  After a U_END that is the first opcode of
    // the copy the next iteration runs.
    for (size_t i = 0; i < Wd.size(); ++i) {
        const UI &u = Wd[i];
        uint32_t next = 0;
        if (u.uop == U_END) next = Wd[u.a / UINSTR_BYTES].uop;
        else if (i + 1 < Wd.size()) next = Wd[i + 1].uop;
        uint64_t *w = &out.image[k_words + i * UINSTR_WORDS];
        w[0] = (uint64_t)next | ((uint64_t)(uint32_t)u.d << 32);
        w[1] = (uint64_t)u.a | ((uint64_t)u.b << 32);
        w[2] = (uint64_t)u.c | ((uint64_t)u.imm << 32);
        w[3] = (uint64_t)u.kinds | ((uint64_t)u.uop << 32);
    }
    out.n_instr = (uint32_t)Wd.size();
    out.code_off = (uint32_t)(k_words * 8);
    out.code_bytes = (uint32_t)(Wd.size() * UINSTR_BYTES);
    out.double_buffered = dbuf;

    // offsets seen from outside: copy A's names (identity map); after an even number of iterations of a
    // double-buffered image the results sit under copy B's names
    std::vector<uint32_t> map_b = slot_map;
    for (uint32_t i = 0; i < dp.n_slots; ++i) slot_map[i] = i;
    out.param_off.assign(dp.param_count, -1);
    for (uint32_t j = 0; j < dp.param_count; ++j)
        if (dp.param_slot[j] != NO_SLOT) out.param_off[j] = (int32_t)((uint32_t)dp.param_slot[j] * SB);
    for (uint32_t f : dp.result_field) {
        const bool is_k = field_kind(f) == KIND_K;
        out.result_is_k.push_back(is_k ? 1 : 0);
        out.result_off.push_back((int32_t)(is_k ? koff(f) : soff(f)));
        out.result_off_even.push_back((int32_t)(is_k ? koff(f) : (dbuf ? map_b[field_index(f)] : field_index(f)) * SB));
    }
    return SAF_OK;
}

} // namespace saf
