//! zolt_gpu.zig -- the Zig side of the C ABI in include/zolt_gpu.h.
//!
//! Drop this file into the reference tree as `src/gpu/zolt_gpu.zig` and apply `zig/build.zig.patch`; the existing
//! call signatures stay intact (SURVEY 8b):
//!   msm.MSM(F, Fp).compute                 src/msm/mod.zig:355-372          -> msmCompute
//!   HyperKZG.commit / batchCommit / open   src/poly/commitment/mod.zig:239-324, :558-570 -> commit / openProof
//!   HyperKZG.batchOpen                     src/poly/commitment/mod.zig:607-732 -> batchOpen
//!   DoryCommitmentScheme row commitments   src/poly/commitment/dory.zig:646-670, :989-1042 -> doryRowCommitments
//!   batched_sumcheck.SumcheckInstance      src/zkvm/batched_sumcheck.zig:34-74 -> GpuSumcheck.instance()
//!   DensePolynomial.bindLow / bindFirst    src/poly/mod.zig:128-175         -> bindLow / bindFirst
//!   EqPolynomial.evals                     src/poly/mod.zig:240-290         -> eqEvals
//! The Keccak / Blake2b transcripts stay on the host: a round hands back (degree + 1) x 32 bytes of Montgomery limbs
//! and takes one 32-byte challenge.
//!
//! NOTE: no Zig toolchain exists in the image this repository is built in, so this file has never been compiled;
//! it is the binding a maintainer would start from.  The same host logic is implemented and tested in C++
//! (zolt_b200/host/zolt_host.hpp), which the parity tests drive.
const std = @import("std");
const c = @cImport({
    @cInclude("zolt_gpu.h");
});
const field = @import("../field/mod.zig");
const msm = @import("../msm/mod.zig");
const batched_sumcheck = @import("../zkvm/batched_sumcheck.zig");

pub const F = field.BN254Scalar; // limbs: [4]u64, Montgomery, canonical == zolt_fr
pub const Fp = field.BN254BaseField; // == zolt_fq
pub const Affine = msm.AffinePoint(Fp);

var g_ctx: ?*c.zolt_ctx = null;

/// One context (device 0, private stream) per process; there is no CPU fallback.
pub fn ctx() *c.zolt_ctx {
    if (g_ctx == null and c.zolt_gpu_init(0, null, &g_ctx) != c.ZOLT_OK)
        @panic("zolt_gpu: no CUDA device (there is no CPU fallback)");
    return g_ctx.?;
}

fn must(st: c_int, what: []const u8) void {
    if (st != c.ZOLT_OK) std.debug.panic("zolt_gpu: {s} -> {d}", .{ what, st });
}

fn pointFrom(xy: [8]u64, inf: u8) Affine {
    if (inf != 0) return Affine.identity(); // (0, 0, infinity = true), src/msm/mod.zig:24-30
    return Affine.fromCoords(.{ .limbs = xy[0..4].* }, .{ .limbs = xy[4..8].* });
}

/// Long-lived device copy of an SRS (SetupParams.powers_of_tau_g1, DorySRS.g1_vec).  AffinePoint is not an extern
/// struct (x, y, infinity: bool with unspecified layout), so it is marshalled once into packed xy[8] + infinity bytes.
pub const DeviceBases = struct {
    handle: *c.zolt_bases,

    pub fn upload(alloc: std.mem.Allocator, pts: []const Affine) !DeviceBases {
        const xy = try alloc.alloc(u64, pts.len * 8);
        defer alloc.free(xy);
        const inf = try alloc.alloc(u8, pts.len);
        defer alloc.free(inf);
        for (pts, 0..) |p, i| {
            @memcpy(xy[8 * i ..][0..4], &p.x.limbs);
            @memcpy(xy[8 * i + 4 ..][0..4], &p.y.limbs);
            inf[i] = @intFromBool(p.infinity);
        }
        var h: ?*c.zolt_bases = null;
        must(c.zolt_msm_bases_upload(ctx(), xy.ptr, inf.ptr, pts.len, &h), "bases_upload");
        return .{ .handle = h.? };
    }
    /// HyperKZG.setup's mock SRS (tau = 0x12345678) generated on the device (commitment/mod.zig:174-213)
    pub fn hyperkzgSetup(max_degree: usize) DeviceBases {
        var h: ?*c.zolt_bases = null;
        must(c.zolt_hyperkzg_setup(ctx(), max_degree, &h), "hyperkzg_setup");
        return .{ .handle = h.? };
    }
    /// 64-byte points straight off a file: fmt = c.ZOLT_G1_WIRE_BE (raw binary SRS), _LE (PTAU tauG1), _ARK_LE (Jolt Dory SRS)
    pub fn fromWire(bytes: []const u8, n: usize, fmt: c_int) !DeviceBases {
        var h: ?*c.zolt_bases = null;
        var bad: usize = 0;
        if (c.zolt_msm_bases_upload_wire(ctx(), bytes.ptr, n, fmt, &h, &bad) != c.ZOLT_OK) return error.PointNotOnCurve;
        return .{ .handle = h.? };
    }
    /// once per SRS: window-level table, 13 instead of 16 mixed additions per point at c = 20
    pub fn precompute(self: DeviceBases) void {
        must(c.zolt_msm_bases_precompute(self.handle, 0), "bases_precompute");
    }
    pub fn deinit(self: DeviceBases) void {
        _ = c.zolt_msm_bases_free(self.handle);
    }
};

/// Body for MSM(F, Fp).compute when the bases are resident (infallible, like the reference).
pub fn msmCompute(bases: DeviceBases, offset: usize, scalars: []const F) Affine {
    var xy: [8]u64 = undefined;
    var inf: u8 = 0;
    must(c.zolt_msm(ctx(), bases.handle, offset, @ptrCast(scalars.ptr), scalars.len, &xy, &inf), "msm");
    return pointFrom(xy, inf);
}

/// HyperKZG.commit: n = min(len(evals), len(powers)); empty -> identity (commitment/mod.zig:239-255)
pub fn commit(powers: DeviceBases, evals: []const F) Affine {
    var len: usize = 0;
    must(c.zolt_msm_bases_len(powers.handle, &len), "bases_len");
    const n = @min(evals.len, len);
    if (n == 0) return Affine.identity();
    return msmCompute(powers, 0, evals[0..n]);
}

/// HyperKZG.open (commitment/mod.zig:261-324): quotient commitments + final_eval, all levels on the device.
pub fn openProof(alloc: std.mem.Allocator, powers: DeviceBases, evals: []const F, point: []const F) !struct { quotients: []Affine, final_eval: F } {
    const nv = point.len;
    const xy = try alloc.alloc(u64, @max(nv, 1) * 8);
    defer alloc.free(xy);
    const inf = try alloc.alloc(u8, @max(nv, 1));
    defer alloc.free(inf);
    var nq: c_int = 0;
    var fin: F = F.zero();
    must(c.zolt_hyperkzg_open(ctx(), powers.handle, @ptrCast(evals.ptr), evals.len, @ptrCast(point.ptr), @intCast(nv), xy.ptr, inf.ptr, &nq, &fin.limbs), "hyperkzg_open");
    const q = try alloc.alloc(Affine, @intCast(nq));
    for (q, 0..) |*p, i| p.* = pointFrom(xy[8 * i ..][0..8].*, inf[i]);
    return .{ .quotients = q, .final_eval = fin };
}

/// The G1 half of DoryCommitmentScheme.commit (dory.zig:989-1042): 2^nu row MSMs over the shared g1_vec, one call.
pub fn doryRowCommitments(alloc: std.mem.Allocator, g1_vec: DeviceBases, evals: []const F) ![]Affine {
    var rows: usize = 0;
    var cols: usize = 0;
    const cap = @max(@as(usize, 1), evals.len);
    const xy = try alloc.alloc(u64, cap * 8);
    defer alloc.free(xy);
    const inf = try alloc.alloc(u8, cap);
    defer alloc.free(inf);
    must(c.zolt_dory_row_commitments(ctx(), g1_vec.handle, @ptrCast(evals.ptr), evals.len, xy.ptr, inf.ptr, &rows, &cols), "dory_rows");
    const out = try alloc.alloc(Affine, rows);
    for (out, 0..) |*p, i| p.* = pointFrom(xy[8 * i ..][0..8].*, inf[i]);
    return out;
}

/// One sumcheck instance on the device behind the reference's SumcheckInstance vtable.
pub const GpuSumcheck = struct {
    sc: *c.zolt_sc,
    num_rounds: usize,
    degree: usize,
    input_claim: F,

    /// tables: k slices of 2^log_n elements each (copied to the device once); combiner: c.ZOLT_SC_*
    pub fn init(combiner: c_int, pairing: c_int, log_n: usize, tables: []const []const F, input_claim: F, degree: usize) GpuSumcheck {
        var ptrs: [8][*c]const u64 = undefined;
        for (tables, 0..) |t, i| ptrs[i] = @ptrCast(t.ptr);
        var h: ?*c.zolt_sc = null;
        must(c.zolt_sc_create(ctx(), combiner, pairing, @intCast(log_n), @intCast(tables.len), 0, &ptrs, &h), "sc_create");
        must(c.zolt_sc_set_claim(h, &input_claim.limbs), "sc_set_claim");
        return .{ .sc = h.?, .num_rounds = log_n, .degree = degree, .input_claim = input_claim };
    }
    fn roundPoly(p: *anyopaque, round: usize) anyerror![4]F {
        _ = round;
        const self: *GpuSumcheck = @ptrCast(@alignCast(p));
        var out: [4]F = .{ F.zero(), F.zero(), F.zero(), F.zero() };
        must(c.zolt_sc_round_evals(self.sc, @ptrCast(&out)), "sc_round_evals"); // evaluations at X = 0, 1, 2, 3
        return out;
    }
    fn bind(p: *anyopaque, challenge: F) anyerror!void {
        const self: *GpuSumcheck = @ptrCast(@alignCast(p));
        must(c.zolt_sc_bind(self.sc, &challenge.limbs), "sc_bind"); // fused with the next round's evaluation
    }
    fn cacheOpenings(p: *anyopaque, r: []const F) anyerror!void {
        _ = p;
        _ = r;
    }
    pub fn instance(self: *GpuSumcheck) batched_sumcheck.SumcheckInstance(F) {
        return .{
            .num_rounds = self.num_rounds,
            .degree = self.degree,
            .input_claim = self.input_claim,
            .computeRoundPolyFn = roundPoly,
            .bindChallengeFn = bind,
            .cacheOpeningsFn = cacheOpenings,
            .ctx = self,
        };
    }
    /// the fully bound value of table t (the opening claim the stage caches)
    pub fn finalValue(self: *GpuSumcheck, t: usize) F {
        var out: F = F.zero();
        must(c.zolt_sc_download(self.sc, @intCast(t), &out.limbs), "sc_download");
        return out;
    }
    pub fn deinit(self: *GpuSumcheck) void {
        _ = c.zolt_sc_free(self.sc);
    }
};

/// One process, several devices (`zolt prove` is a single process): contexts + connected mailboxes for device_ids.
pub fn initMulti(device_ids: []const c_int) *c.zolt_multi {
    var m: ?*c.zolt_multi = null;
    must(c.zolt_gpu_init_multi(device_ids.ptr, @intCast(device_ids.len), &m), "gpu_init_multi");
    return m.?;
}

/// Stage 3 building blocks over device tables (src/zkvm/spartan/stage3_prover.zig).  `d_*` are device pointers obtained
/// from zolt_dev_alloc / zolt_dev_upload; the sequence of calls for the three provers is zolt_host.hpp `stage3::*`.
/// eq+1(r, .) as a table (computeEqPlusOneEvals, :1878-1895); r big-endian like EqPolynomial.evals.
pub fn eqPlusOneEvals(r: []const F, d_out: *anyopaque) void {
    must(c.zolt_eq_plus_one_evals(ctx(), @ptrCast(r.ptr), @intCast(r.len), d_out), "eq_plus_one_evals");
}
/// Q[x_lo] = scale * sum_{x_hi} v[x_lo + x_hi * n_lo] * suffix[x_hi]  (ShiftPrefixSuffixProver.init :1053-1112,
/// RegistersPrefixSuffixProver.init :2262-2300)
pub fn foldHigh(d_q: *anyopaque, d_v: *const anyopaque, d_suffix: *const anyopaque, n_lo: usize, n_hi: usize, scale: ?*const F) void {
    must(c.zolt_fr_fold_high_dev(ctx(), d_q, d_v, d_suffix, n_lo, n_hi, if (scale) |s| &s.limbs else null), "fr_fold_high_dev");
}

/// LassoProver address round (src/zkvm/lasso/prover.zig:265-310, :380-407): applies the previous challenge (bind_bit < 0:
/// none) and returns [sum_0, sum_1] for bit `sum_bit` of the lookup indices in one pass over eq_evals.
/// lookup_indices is the prover's `[]const u128` uploaded as is (16 bytes each, little-endian).
pub fn lassoAddrRound(d_eq_evals: *anyopaque, d_indices: *const anyopaque, n: usize, bind_bit: c_int, r: ?*const F, sum_bit: c_int) [2]F {
    var sums: [8]u64 = undefined;
    must(c.zolt_lasso_addr_round(ctx(), d_eq_evals, d_indices, n, bind_bit, if (r) |x| &x.limbs else null, sum_bit, &sums), "lasso_addr_round");
    return .{ .{ .limbs = sums[0..4].* }, .{ .limbs = sums[4..8].* } };
}
