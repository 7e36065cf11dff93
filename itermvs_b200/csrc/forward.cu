// IterMVS.forward in test mode (reference models/itermvs.py:253-329) as one host call that
// enqueues every kernel of the hot path on the caller's stream.  No allocation, no sync: the
// caller provides the workspace, so the whole call is CUDA-graph capturable.
#include "common.cuh"

namespace imvs {
int tune(const char* name, int def);   // defined in warp.cu
int conv_passes();
int forward_prologue(const float* proj1, const float* proj2, const float* proj3, int B, int V, float* rt1, float* rt2, float* rt3,
                     int* nan_flag, const float* fea3, float* fea3p, int H3, int W3, cudaStream_t st);      // warp.cu
}

namespace imvs {

struct Workspace {
    float *rt1, *rt2, *rt3;
    float *corr_init, *pvw_logits, *vw3, *vw2, *agg_init, *corrnet_scratch, *corr0;
    float *hinit_scratch, *hidden, *xbuf, *agg_iter, *gru_scratch, *head_scratch, *ups_scratch, *conf_buf;
    float *fea3p;          // level-3 pyramid padded to 64 floats per texel (imvs_pad_level3)
    size_t total_floats;
};

static size_t take(size_t& cursor, size_t n) {
    size_t at = cursor;
    cursor += (n + 63) / 64 * 64;      // 256-byte granules keep every buffer float4-aligned
    return at;
}

static Workspace carve(const imvs_problem& pb, float* base) {
    const size_t B = pb.B, S = pb.V - 1, D = pb.D;
    const size_t P2 = (size_t)(pb.H / 4) * (pb.W / 4), P3 = (size_t)(pb.H / 8) * (pb.W / 8);
    size_t c = 0;
    Workspace w;
    auto at = [&](size_t n) { return base + take(c, n); };
    w.rt1 = at(B * S * 12);
    w.rt2 = at(B * S * 12);
    w.rt3 = at(B * S * 12);
    w.corr_init = at(B * S * D * P3 * 8);
    w.pvw_logits = at(B * S * D * P3);
    w.vw3 = at(B * S * P3);
    w.vw2 = at(B * S * P2);
    w.agg_init = at(B * D * P3 * 8);
    size_t cn = B * D * P3 * 48, ci = B * IMVS_ITER_SLICES * P2 * 48;     // = imvs_corrnet_scratch_floats
    w.corrnet_scratch = at(cn > ci ? cn : ci);
    w.corr0 = at(B * D * P3);
    w.hinit_scratch = at(B * 96 * P3);
    w.hidden = at(B * 32 * P2);
    w.xbuf = at(B * IMVS_XCH * P2);
    w.agg_iter = at(B * IMVS_ITER_SLICES * P2 * 8);
    w.gru_scratch = at(B * 128 * P2);
    w.head_scratch = at(B * 384 * P2);
    w.ups_scratch = at(B * 64 * P2);
    w.conf_buf = at(B * P2);
    w.fea3p = at(B * (S + 1) * P3 * 64);
    w.total_floats = c;
    return w;
}

static int check_problem(const imvs_problem* pb) {
    IMVS_REQUIRE(pb, "null problem");
    IMVS_REQUIRE(pb->B >= 1, "B=%d", pb->B);
    IMVS_REQUIRE(pb->V >= 2 && pb->V - 1 <= IMVS_MAX_VIEWS, "need 1..%d source views (V=%d)", IMVS_MAX_VIEWS, pb->V);
    IMVS_REQUIRE(pb->H >= 32 && pb->W >= 32 && pb->H % 32 == 0 && pb->W % 32 == 0,
                 "H and W must be multiples of 32 (H=%d W=%d): CorrNet halves the 1/8-resolution map twice", pb->H, pb->W);
    IMVS_REQUIRE(pb->D >= 8 && pb->D % 8 == 0, "D=%d must be a multiple of 8", pb->D);
    IMVS_REQUIRE(pb->iterations >= 1, "iterations=%d", pb->iterations);
    return 0;
}

}  // namespace imvs

using namespace imvs;

extern "C" size_t imvs_forward_workspace_bytes(const imvs_problem* pb) {
    if (check_problem(pb) != 0) return 0;
    return carve(*pb, nullptr).total_floats * sizeof(float);
}

extern "C" int imvs_forward_launch_count(const imvs_problem* pb) {
    if (check_problem(pb) != 0) return -1;
    // the depth head is 2 launches (conv0 + the fused tcgen05 kernel) in the default fp32-grade mode, 4 otherwise
    // the ConvGRU is 3 launches (operand split + z|r + q on the TMA / tcgen05 kernel) in the default mode, 2 otherwise
#ifdef CUSIM
    const int head = 4, gru = 2;
#else
    const int head = (conv_passes() == 4 && tune("HEADFUSED", 1)) ? 2 : 4;
    const int gru = (conv_passes() == 4 && tune("TC5P_GRU", 1)) ? 3 : 2;
#endif
    // CorrNet is six launches per pass (one with IMVS_TUNE_CORR_TILE=1 / CORR_FUSED=1, seven with TC5P_CORR=1: experiment switches)
    const int corr = tune("CORR_TILE", 0) ? 1 : 6;
    // default: ONE prologue launch (three projection compositions + the padded level-3 copy); IMVS_TUNE_WC_PAD3=0: three compose launches
    return (tune("WC_PAD3", 1) ? 12 : 14) + corr + head + (1 + corr + gru + head) * pb->iterations;
}

extern "C" int imvs_itermvs_forward(const imvs_problem* pb, const imvs_weights* w,
                                    const float* fea1, const float* fea2, const float* fea3,
                                    const float* proj1, const float* proj2, const float* proj3,
                                    const float* depth_min, const float* depth_max,
                                    void* workspace, size_t workspace_bytes,
                                    float* depth, float* depth_up, float* conf, float* conf_up,
                                    int* nan_flag, void* stream) {
    IMVS_TRY(check_problem(pb));
    IMVS_REQUIRE(w && fea1 && fea2 && fea3 && proj1 && proj2 && proj3 && depth_min && depth_max && workspace,
                 "itermvs_forward: null pointer");
    IMVS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "itermvs_forward: workspace must be 256-byte aligned");
    ApiScope api_;
    Workspace ws = carve(*pb, static_cast<float*>(workspace));
    IMVS_REQUIRE(workspace_bytes >= ws.total_floats * sizeof(float), "itermvs_forward: workspace too small (%zu < %zu bytes)",
                 workspace_bytes, ws.total_floats * sizeof(float));
    const int B = pb->B, V = pb->V, S = V - 1, D = pb->D, I = pb->iterations;
    const int H2 = pb->H / 4, W2 = pb->W / 4, H3 = pb->H / 8, W3 = pb->W / 8;
    const int P2 = H2 * W2, P3 = H3 * W3;
    const size_t xb = (size_t)IMVS_XCH * P2, xp = IMVS_XCH;      // x: [B][P2][16], ch 0 = normalized depth

    // x channels 11..15 are zero padding of the GRU's 43 -> 48 input channels
    IMVS_CUDA(cudaMemsetAsync(ws.xbuf, 0, sizeof(float) * (size_t)B * xb, (cudaStream_t)stream));

    // default (IMVS_TUNE_WC_PAD3=0 switches it off): the plane-sweep kernels read level 3 from a copy padded to 256 bytes per texel
    // (one full line + one 64-byte piece per tap and lane group instead of three half-used 64-byte pieces, a lane owns its correlation
    // group: no regrouping shuffles; warpcorr.cu), written by the same launch that composes the projections
    const bool pad3 = tune("WC_PAD3", 1) != 0;
    // K1 (module.py:78-90), hoisted: once per level instead of once per warp call
    { StageTimer tm_(ST_COMPOSE, stream);
    if (pad3) {
        IMVS_TRY(forward_prologue(proj1, proj2, proj3, B, V, ws.rt1, ws.rt2, ws.rt3, nan_flag, fea3, ws.fea3p, H3, W3, (cudaStream_t)stream));
    } else {
        IMVS_TRY(imvs_compose_projections(proj1, B, V, ws.rt1, nan_flag, stream));
        IMVS_TRY(imvs_compose_projections(proj2, B, V, ws.rt2, nan_flag, stream));
        IMVS_TRY(imvs_compose_projections(proj3, B, V, ws.rt3, nan_flag, stream));
    }
    }

    // init evaluation (itermvs.py:270-271, 36-70)
    { StageTimer tm_(ST_WARPCORR_INIT, stream);
    if (pad3) {
        IMVS_TRY(imvs_warpcorr_init_padded(ws.fea3p, ws.rt3, depth_min, depth_max, nullptr, ws.corr_init, B, V, H3, W3, D, stream));
    } else {
        IMVS_TRY(imvs_warpcorr_init(fea3, ws.rt3, depth_min, depth_max, nullptr, ws.corr_init, B, V, H3, W3, D, stream));
    }
    }
    { StageTimer tm_(ST_PVW, stream);
    IMVS_TRY(imvs_pixel_view_weight(w, ws.corr_init, ws.pvw_logits, ws.vw3, ws.vw2, B, S, D, H3, W3, stream));
    }
    { StageTimer tm_(ST_AGG_INIT, stream);
    IMVS_TRY(imvs_aggregate_init(ws.corr_init, ws.vw3, ws.agg_init, B, S, D, P3, stream));
    }
    imvs_corrnet_weights init_sets[3] = {w->corrnet[2], w->corrnet[2], w->corrnet[2]};
    { StageTimer tm_(ST_CORRNET, stream);       // -> corr0 [B][P3][D] channels-last
    IMVS_TRY(imvs_corrnet(init_sets, D, D, D, ws.agg_init, ws.corr0, (size_t)D * P3, (size_t)D, ws.corrnet_scratch, B * D, H3, W3, stream));
    }

    // hidden state and first depth (itermvs.py:275-276)
    { StageTimer tm_(ST_HIDDEN_INIT, stream);
    IMVS_TRY(imvs_hidden_init(w, ws.corr0, ws.hidden, ws.hinit_scratch, B, D, H3, W3, stream));
    }
    { StageTimer tm_(ST_HEAD, stream);
    IMVS_TRY(imvs_depth_head(w, ws.hidden, ws.xbuf, xb, xp, nullptr, nullptr, nullptr, (I == 1) ? depth : nullptr,
                             depth_min, depth_max, ws.head_scratch, B, H2, W2, stream));
    }

    float* conf_q = conf ? conf : ws.conf_buf;
    for (int it = 0; it < I; ++it) {
        const bool last = (it == I - 1);
        // itermvs.py:288-295
        { StageTimer tm_(ST_WARPCORR_ITER, stream);
        if (pad3)
            IMVS_TRY(imvs_warpcorr_iter_padded(fea1, fea2, ws.fea3p, ws.rt1, ws.rt2, ws.rt3, ws.xbuf, xb, xp, ws.vw2, depth_min, depth_max,
                                               nullptr, nullptr, nullptr, ws.agg_iter, B, V, H2, W2, stream));
        else
            IMVS_TRY(imvs_warpcorr_iter(fea1, fea2, fea3, ws.rt1, ws.rt2, ws.rt3, ws.xbuf, xb, xp, ws.vw2, depth_min, depth_max,
                                        nullptr, nullptr, nullptr, ws.agg_iter, B, V, H2, W2, stream));
        }
        { StageTimer tm_(ST_CORRNET, stream);   // -> x channels 1..10
        IMVS_TRY(imvs_corrnet(w->corrnet, IMVS_ITER_SLICES, 4, 8, ws.agg_iter, ws.xbuf + 1, xb, xp, ws.corrnet_scratch,
                              B * IMVS_ITER_SLICES, H2, W2, stream));
        }
        // itermvs.py:316-320 -> Update.forward (192-220); x = [normalized_depth, corr] is xbuf itself
        { StageTimer tm_(ST_GRU, stream);
        IMVS_TRY(imvs_conv_gru(w, ws.hidden, ws.xbuf, ws.gru_scratch, B, H2, W2, stream));
        }
        // `depth` returned in test mode is the value BEFORE the last update (itermvs.py:319)
        float* depth_here = (!last && it == I - 2) ? depth : nullptr;
        { StageTimer tm_(ST_HEAD, stream);
        IMVS_TRY(imvs_depth_head(w, ws.hidden, ws.xbuf, xb, xp, nullptr, last ? conf_q : nullptr, nullptr, depth_here,
                                 depth_min, depth_max, ws.head_scratch, B, H2, W2, stream));
        }
    }
    // itermvs.py:321-324
    if (depth_up || conf_up) {
        IMVS_REQUIRE(depth_up, "itermvs_forward: conf_up requested without depth_up");
        StageTimer tm_(ST_UPSAMPLE, stream);
        IMVS_TRY(imvs_upsample_outputs(w, fea2, (size_t)V * P2 * 32, ws.xbuf, xb, xp, conf_up ? conf_q : nullptr, depth_min, depth_max,
                                       depth_up, conf_up, ws.ups_scratch, B, H2, W2, stream));
    }
    return 0;
}
