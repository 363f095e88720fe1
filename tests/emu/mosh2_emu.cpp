// mosh2_emu.cpp -- TEST-ONLY host build of the CTA program (moshpp_b200/csrc/mosh2_device.cuh).
//
// Compiles the device source with MOSH2_EMU: one "thread" (tid 0 of 1), barriers are no-ops.  The CPU
// test-suite runs it against the oracle to validate index math and control flow of the exact code that
// nvcc compiles for sm_100a.  It is never linked into libmosh2.so and never loaded by moshpp_b200.
#define MOSH2_EMU 1
#include "../../include/mosh2.h"
#include "../../moshpp_b200/csrc/mosh2_device.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

template <class real>
struct HostModel {
    mosh2::Model<real> m{};
    std::vector<std::vector<char>> store;
    template <class T, class U>
    const T *up(const U *src, size_t n) {
        store.emplace_back((n + 1) * sizeof(T));
        T *p = reinterpret_cast<T *>(store.back().data());
        for (size_t i = 0; i < n; ++i) p[i] = static_cast<T>(src[i]);
        return p;
    }
    void build(const mosh2_model_desc &d) {
        const size_t nJ = d.n_joints, S = size_t(3) * d.n_markers, nd = d.n_dmpl;
        m.nJ = d.n_joints; m.M = d.n_markers; m.body_dof = d.body_dof; m.p_red = d.p_red;
        m.n_hand_red = d.n_hand_red; m.n_hand_full = d.n_hand_full; m.nd = d.n_dmpl; m.kw = d.kw; m.na = d.na;
        m.n_levels = d.n_levels; m.prior_k = d.prior_k; m.prior_d = d.prior_d; m.prior_off = d.prior_off;
        m.n1 = d.n_free1; m.n2 = d.n_free2; m.finger_lo = d.finger_lo; m.finger_hi = d.finger_hi;
        m.parents = up<int>(d.parents, nJ);
        m.fk_order = up<int>(d.fk_order, nJ);
        m.level_ofs = up<int>(d.level_ofs, size_t(d.n_levels) + 1);
        m.w_joint = up<int>(d.w_joint, S * d.kw);
        m.anc_joint = up<int>(d.anc_joint, S * d.na);
        m.anc_mask = up<int>(d.anc_mask, S * d.na);
        m.anc_pos = up<int8_t>(d.anc_pos, S * nJ);
        std::vector<int> lo(d.n_hand_red, 0), hi(d.n_hand_red, 0);
        for (int r = 0; r < d.n_hand_red; ++r) {
            int a = d.n_hand_full, b = 0;
            for (int c = 0; c < d.n_hand_full; ++c)
                if (d.hand_comps[size_t(r) * d.n_hand_full + c] != 0.0) { if (c < a) a = c; b = c + 1; }
            if (b <= a) a = b = 0;
            lo[r] = a; hi[r] = b;
        }
        m.hand_lo = up<int>(lo.data(), lo.size());
        m.hand_hi = up<int>(hi.data(), hi.size());
        m.hand_comps = up<real>(d.hand_comps, size_t(d.n_hand_red) * d.n_hand_full);
        m.hands_mean = up<real>(d.hands_mean, d.n_hand_full);
        m.v0 = up<real>(d.v0, S * 3);
        m.sd = up<real>(d.sd, S * 3 * nd);
        m.pd = up<real>(d.pd, (nJ - 1) * S * 3 * 9);
        m.w_val = up<real>(d.w_val, S * d.kw);
        m.j0 = up<real>(d.j0, nJ * 3);
        m.jd = up<real>(d.jd, nJ * 3 * nd);
        m.coefs = up<real>(d.coefs, size_t(d.n_markers) * 3);
        m.prior_means = up<real>(d.prior_means, size_t(d.prior_k) * d.prior_d);
        m.prior_Q = up<real>(d.prior_Q, size_t(d.prior_k) * d.prior_d * d.prior_d);
        m.prior_nlw = up<real>(d.prior_neglogw, d.prior_k);
        m.free1 = up<int>(d.free1, d.n_free1);
        m.free2 = up<int>(d.free2, d.n_free2);
    }
};

template <class real>
int run(const mosh2_model_desc *desc, const mosh2_options *opt, int n_frames, const double *obs, const uint8_t *vis,
        int chunk_len, int warmup, const mosh2_result *res) {
    HostModel<real> hm;
    hm.build(*desc);
    const size_t F = n_frames, M = desc->n_markers, PF = size_t(3) * desc->n_joints, PR = desc->p_red, nd = desc->n_dmpl;
    std::vector<real> o(F * M * 3);
    for (size_t i = 0; i < o.size(); ++i) o[i] = real(obs[i]);
    std::vector<real> fullpose(F * PF), pose(F * PR), trans(F * 3), dmpls(F * nd + 1), mk(F * M * 3), errs(F * 6);
    mosh2::Job<real> job{};
    job.n_frames = n_frames;
    job.chunk_len = chunk_len > 0 ? chunk_len : 0;
    job.warmup = warmup > 0 ? warmup : 0;
    job.n_chunks = job.chunk_len ? (n_frames + job.chunk_len - 1) / job.chunk_len : 1;
    job.obs = o.data(); job.vis = vis;
    job.fullpose = fullpose.data(); job.pose = pose.data(); job.trans = trans.data();
    job.dmpls = nd ? dmpls.data() : nullptr; job.markers_sim = mk.data(); job.errs = errs.data();
    std::vector<int> status(F, 0), counters(F * 4, 0);
    job.status = status.data(); job.counters = counters.data();
    int totals[4] = {0, 0, 0, 0};
    job.totals = totals;
    job.gws = nullptr; job.gws_stride = 0;
    mosh2::Options &q = job.opt;
    q.wt_data = opt->wt_data; q.wt_poseB = opt->wt_poseB; q.wt_poseH = opt->wt_poseH; q.wt_velo = opt->wt_velo;
    q.wt_dmpl = opt->wt_dmpl; q.wt_annealing = opt->wt_annealing; q.wt_extrap = opt->wt_extrap_dmpl;
    q.num_train_markers = opt->num_train_markers; q.delta_0 = opt->delta_0; q.e3_first = opt->e3_first; q.e3 = opt->e3;
    q.maxiter = opt->maxiter; q.optimize_fingers = opt->optimize_fingers; q.optimize_dynamics = opt->optimize_dynamics;

    const mosh2::Dims d = mosh2::make_dims(hm.m);
    mosh2::Work<real> w;
    mosh2::Arena S0{nullptr, 0}, G0{nullptr, 0};
    mosh2::carve(w, d, S0, G0, false);
    std::vector<char> smem(S0.off + 64);
    for (int c = 0; c < job.n_chunks; ++c) {
        std::memset(smem.data(), 0, smem.size());
        mosh2::Arena S{smem.data(), 0}, G{nullptr, 0};
        mosh2::carve(w, d, S, G, false);
        mosh2::Cta cta{0, 1};
        mosh2::Solver<real> s(hm.m, job, w, cta);
        s.run_chunk(c);
    }
    auto conv = [](double *dst, const std::vector<real> &src, size_t n) {
        if (dst) for (size_t i = 0; i < n; ++i) dst[i] = double(src[i]);
    };
    conv(res->fullpose, fullpose, F * PF);
    conv(res->pose, pose, F * PR);
    conv(res->trans, trans, F * 3);
    if (nd) conv(res->dmpls, dmpls, F * nd);
    conv(res->markers_sim, mk, F * M * 3);
    conv(res->errs, errs, F * 6);
    if (res->status) std::memcpy(res->status, status.data(), F * sizeof(int));
    if (res->counters) std::memcpy(res->counters, counters.data(), F * 4 * sizeof(int));
    return 0;
}

}  // namespace

extern "C" int mosh2_emu_solve(const mosh2_model_desc *desc, const mosh2_options *opt, int32_t n_frames,
                               const double *obs, const uint8_t *vis, int32_t chunk_len, int32_t warmup,
                               int32_t precision, const mosh2_result *res) {
    if (precision == MOSH2_F64) return run<double>(desc, opt, n_frames, obs, vis, chunk_len, warmup, res);
    return run<float>(desc, opt, n_frames, obs, vis, chunk_len, warmup, res);
}
