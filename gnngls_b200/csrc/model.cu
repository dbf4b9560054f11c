// gnngls_model_forward: EdgePropertyPredictionModel.forward (gnngls/models.py:65-70) for a batch of line graphs of K_n as
// one C-ABI call -- the per-op entry points of this library issued back to back on the caller's stream.  The workspace is
// carved into the activation ping-pong buffers (fp32 residual stream + operand copies), the projected features, the
// attention scores and the two kernels' own workspaces.
#include <cstdint>
#include "common.h"

namespace {
constexpr size_t ALIGN = 256;
size_t up(size_t x) { return (x + ALIGN - 1) / ALIGN * ALIGN; }

struct Carve {
    size_t ha, hb, ha_op, hb_op, ft, el, er, h1, gat, ff, total;
    Carve(int B, int n, int dense_impl) {
        const size_t M = (size_t)B * ((size_t)n * (n - 1) / 2);
        const bool tc = dense_impl != GNNGLS_DENSE_SIMT;
        size_t o = 0;
        ha = o; o += up(M * 128 * 4);
        hb = o; o += up(M * 128 * 4);
        ha_op = o; o += tc ? up(M * 128 * 4) : 0;            // (fp16 copies use half of it)
        hb_op = o; o += tc ? up(M * 128 * 4) : 0;
        ft = o; o += up(M * 128 * 4);
        el = o; o += up(M * 8 * 4);
        er = o; o += up(M * 8 * 4);
        h1 = o; o += up(M * 128 * 4);
        gat = o; o += up(gnngls_gat_kn_workspace_bytes(B, n));
        ff = o; o += up(gnngls_ff_workspace_bytes(dense_impl, (int64_t)M));
        total = o;
    }
};
}  // namespace

extern "C" size_t gnngls_sizeof_model_args(void) { return sizeof(gnngls_model_args); }

extern "C" size_t gnngls_model_forward_workspace_bytes(int B, int n, int dense_impl) {
    if (B <= 0 || n < 3) return 0;
    return Carve(B, n, dense_impl).total;
}

extern "C" int gnngls_model_forward(const gnngls_model_args *a, void *workspace, size_t workspace_bytes, void *stream) {
    GNNGLS_REQUIRE(a && a->x && a->We && a->be && a->Wd && a->bd && a->layers && a->y, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(a->n >= 3 && a->in_dim >= 1 && a->out_dim >= 1 && a->n_layers >= 0, GNNGLS_ERR_BAD_ARG, "bad dimensions");
    GNNGLS_REQUIRE(a->dense_impl == GNNGLS_DENSE_TCGEN05 || a->dense_impl == GNNGLS_DENSE_SIMT || a->dense_impl == GNNGLS_DENSE_TCGEN05_F16,
                   GNNGLS_ERR_BAD_ARG, "unknown dense_impl %d", a->dense_impl);
    if (a->B <= 0) return GNNGLS_OK;
    const Carve c(a->B, a->n, a->dense_impl);
    GNNGLS_REQUIRE(workspace && workspace_bytes >= c.total, GNNGLS_ERR_WORKSPACE, "model workspace too small: need %zu bytes", c.total);
    GNNGLS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 127) == 0, GNNGLS_ERR_BAD_ARG, "model workspace must be 128-byte aligned");
    unsigned char *w = static_cast<unsigned char *>(workspace);
    const int64_t M = (int64_t)a->B * ((int64_t)a->n * (a->n - 1) / 2);
    const bool tc = a->dense_impl != GNNGLS_DENSE_SIMT;
    const int op_dtype = a->dense_impl == GNNGLS_DENSE_TCGEN05_F16 ? GNNGLS_FT_F16 : GNNGLS_FT_TF32;
    float *cur = reinterpret_cast<float *>(w + c.ha), *nxt = reinterpret_cast<float *>(w + c.hb);
    void *cur_op = tc ? w + c.ha_op : nullptr, *nxt_op = tc ? w + c.hb_op : nullptr;
    void *ft = w + c.ft;
    float *el = reinterpret_cast<float *>(w + c.el), *er = reinterpret_cast<float *>(w + c.er), *h1 = reinterpret_cast<float *>(w + c.h1);
    int rc = gnngls_embed_forward(a->x, M, a->in_dim, a->We, a->be, cur, cur_op, op_dtype, stream);
    if (rc != GNNGLS_OK) return rc;
    for (int l = 0; l < a->n_layers; ++l) {
        const gnngls_layer_params &p = a->layers[l];
        GNNGLS_REQUIRE(p.Wfc && p.attn_l && p.attn_r && p.bn1_scale && p.bn1_shift && p.W1 && p.b1 && p.W2 && p.b2 && p.bn2_scale && p.bn2_shift,
                       GNNGLS_ERR_BAD_ARG, "layer %d: null parameter", l);
        rc = gnngls_fc_forward(a->dense_impl, tc ? cur_op : static_cast<void *>(cur), M, p.Wfc, p.attn_l, p.attn_r, ft, a->ft_dtype, el, er, stream);
        if (rc != GNNGLS_OK) return rc;
        rc = gnngls_gat_aggregate_kn(a->B, a->n, ft, a->ft_dtype, el, er, cur, p.gat_bias, p.bn1_scale, p.bn1_shift, h1, nullptr,
                                     w + c.gat, c.ff - c.gat, stream);
        if (rc != GNNGLS_OK) return rc;
        rc = gnngls_ff_forward(a->dense_impl, h1, nullptr, M, p.W1, p.b1, p.W2, p.b2, p.bn2_scale, p.bn2_shift, nxt, nxt_op, op_dtype,
                               w + c.ff, c.total - c.ff, stream);
        if (rc != GNNGLS_OK) return rc;
        float *t = cur; cur = nxt; nxt = t;
        void *to = cur_op; cur_op = nxt_op; nxt_op = to;
    }
    return gnngls_decision_forward(cur, M, a->out_dim, a->Wd, a->bd, a->y, stream);
}
