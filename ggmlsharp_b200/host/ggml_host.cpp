// Host-side mirror of the slice of GGMLSharp's public API that sits on the mul_mat path, written in
// C++ because no .NET toolchain exists in this image (the C# bindings a maintainer would add are in
// csharp/ and INTEGRATION.md).  Names, argument meaning, arena rule and error behaviour follow the
// reference so that tests read like Test0/Test3 of the reference:
//
//   ggml_init / ggml_free                     Ggml.cs:1447-1601   (seam A: pool from ggb_pool_alloc)
//   ggml_new_tensor{,_1d,_2d,_3d,_4d}         Ggml.cs:2347-2395, 7722-7866
//   ggml_nelements / ggml_nbytes / ...        Ggml.cs:3751-3795
//   ggml_set_f32 / ggml_get_f32_1d            Ggml.cs:2501-2564, 2802-2835  (F32 tensors only)
//   ggml_mul_mat                              Ggml.cs:7137-7151, 8222-8246
//   ggml_cpy                                  Ggml.cs:7169-7175, 8275-8305
//   ggml_build_forward{,_expand}              Ggml.cs:7559-7673
//   ggml_graph_compute                        Ggml.cs:3209-3736   (seam B: MUL_MAT/CPY nodes -> CUDA stream executor)
//
// Everything arithmetic happens in libggb200.so; this file only builds tensor headers and graphs.
//   ggml_add / mul / silu / rms_norm / scale / repeat / cont / transpose (builders)   Ggml.cs:6846-7225, 7868-8322
// Ops outside the path: ggml_graph_compute reports them and stops.
#include "../../include/ggb200.h"

#include <cstdio>
#include <cstring>
#include <mutex>

extern "C" {

typedef struct ggml_init_params {          // TypeDefinitions.cs:24-30
    uint64_t mem_size;
    void    *mem_buffer;
    uint8_t  no_alloc;
} ggml_init_params;

typedef struct ggml_object {               // TypeDefinitions.cs:48-56
    uint64_t offs, size;
    struct ggml_object *next;
    uint8_t padding[8];
} ggml_object;

typedef struct ggml_scratch { uint64_t offs, size; void *data; } ggml_scratch;

typedef struct ggml_context {              // TypeDefinitions.cs:32-46 (88 bytes)
    uint64_t mem_size;
    void    *mem_buffer;
    uint8_t  mem_buffer_owned;
    uint8_t  no_alloc;
    int32_t  n_objects;
    ggml_object *objects_begin, *objects_end;
    ggml_scratch scratch, scratch_save;
} ggml_context;

} // extern "C"

static_assert(sizeof(ggml_object) == GGML_OBJECT_SIZE, "ggml_object");
static_assert(sizeof(ggml_context) == 88, "ggml_context");

namespace {

constexpr int GGML_MAX_CONTEXTS = 64;
constexpr int GGML_DEFAULT_N_THREADS = 4;      // Ggml.cs:22
constexpr int CACHE_LINE_SIZE = 64;
constexpr int GGML_OP_SQR_ = 6;                // TypeDefinitions.cs:180 -- an op the backend does not take

struct Slot { bool used; ggml_context ctx; ggb_pool *pool; };
Slot g_state[GGML_MAX_CONTEXTS];
bool g_first_call = true;
std::mutex g_cs;                               // ggml_critical_section_start/end (Ggml.cs:8451-8470)
int g_last_status = 0;

const int BLCK[GGML_TYPE_COUNT] = {1, 1, 32, 32, 16, 16, 32, 32, 32, 32, 1, 1, 1};                 // Ggml.cs:55-70
const uint64_t TSIZE[GGML_TYPE_COUNT] = {4, 2, 20, 24, 10, 12, 22, 24, 36, 44, 1, 2, 4};           // Ggml.cs:72-87

ggb_pool *pool_of(const ggml_context *ctx)
{
    for (auto &s : g_state) if (s.used && &s.ctx == ctx) return s.pool;
    return nullptr;
}
// the pool whose arena holds these bytes (ggml_set_* take a tensor, not a context)
ggb_pool *pool_holding(const void *p)
{
    const uint8_t *q = static_cast<const uint8_t *>(p);
    for (auto &s : g_state)
        if (s.used && q >= static_cast<const uint8_t *>(s.ctx.mem_buffer) && q < static_cast<const uint8_t *>(s.ctx.mem_buffer) + s.ctx.mem_size) return s.pool;
    return nullptr;
}

} // namespace

extern "C" {

int ggml_host_last_status(void) { return g_last_status; }
void *ggml_host_pool_of(const ggml_context *ctx) { return pool_of(ctx); }       // tests drive seam B directly with flags
// Opt in to weight residency for a context: leaf src0 tensors stay on the device between ggml_graph_compute calls.
// Off by default, because the reference re-reads src0->data on every compute (Ggml.cs:6139-6164) and user code writes tensor->data directly.
int ggml_host_set_weight_cache(ggml_context *ctx, int on)
{
    ggb_pool *pool = pool_of(ctx);
    return pool ? ggb_pool_set_weight_cache(pool, on) : GGB_E_INVALID;
}

// Opt in to the row split across the GPUs of the box for a context (include/ggb200.h: ggb_pool_set_row_split)
int ggml_host_set_row_split(ggml_context *ctx, int max_devices, size_t min_weight_bytes)
{
    ggb_pool *pool = pool_of(ctx);
    return pool ? ggb_pool_set_row_split(pool, max_devices, min_weight_bytes) : GGB_E_INVALID;
}

ggml_context *ggml_init(ggml_init_params params)
{
    std::lock_guard<std::mutex> lk(g_cs);
    if (g_first_call) {
        // where ggml_init_cublas() sat (Ggml.cs:1499-1504): refuse to run on a layout mismatch, then bring the device up
        g_last_status = ggb_abi_check((int)sizeof(ggml_tensor), (int)offsetof(ggml_tensor, data), (int)sizeof(ggml_cgraph),
                                      (int)offsetof(ggml_cgraph, nodes), (int)sizeof(block_q4_0), (int)sizeof(block_q4_1));
        // the device itself is brought up by the first pool allocation / compute (ggb_init is idempotent)
        if (g_last_status != 0) { fprintf(stderr, "ggml_init: %s\n", ggb_last_error()); return nullptr; }
        for (auto &s : g_state) s.used = false;
        g_first_call = false;
    }
    Slot *slot = nullptr;
    for (auto &s : g_state) if (!s.used) { slot = &s; break; }
    if (!slot) return nullptr;                                   // Ggml.cs:1529-1536

    const uint64_t mem_size = (params.mem_size + GGML_MEM_ALIGN - 1) & ~(uint64_t)(GGML_MEM_ALIGN - 1);
    void *buf = params.mem_buffer;
    ggb_pool *pool = nullptr;
    g_last_status = buf ? ggb_pool_adopt(buf, mem_size, &pool) : ggb_pool_alloc(mem_size, &buf, &pool);   // Ggml.cs:1543-1545
    if (g_last_status != 0) { fprintf(stderr, "ggml_init: %s\n", ggb_last_error()); return nullptr; }
    slot->used = true;
    slot->pool = pool;
    memset(&slot->ctx, 0, sizeof slot->ctx);
    slot->ctx.mem_size = mem_size;
    slot->ctx.mem_buffer = buf;
    slot->ctx.mem_buffer_owned = params.mem_buffer ? 0 : 1;
    slot->ctx.no_alloc = params.no_alloc;
    return &slot->ctx;
}

void ggml_free(ggml_context *ctx)
{
    std::lock_guard<std::mutex> lk(g_cs);
    for (auto &s : g_state)
        if (s.used && &s.ctx == ctx) {
            s.used = false;
            ggb_pool_free(s.pool);                               // frees the host pool only if owned (Ggml.cs:1584-1588)
            s.pool = nullptr;
            return;
        }
}

int64_t ggml_nelements(const ggml_tensor *t) { return t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]; }                 // Ggml.cs:3766
int64_t ggml_nrows(const ggml_tensor *t) { return t->ne[1] * t->ne[2] * t->ne[3]; }
uint64_t ggml_nbytes(const ggml_tensor *t) { return (uint64_t)ggml_nelements(t) * TSIZE[t->type] / (uint64_t)BLCK[t->type]; }   // Ggml.cs:3779
int ggml_blck_size(int type) { return BLCK[type]; }
uint64_t ggml_type_size(int type) { return TSIZE[type]; }
uint64_t ggml_used_mem(const ggml_context *ctx) { return ctx->objects_end ? ctx->objects_end->offs + ctx->objects_end->size : 0; }

// Ggml.cs:7722-7866 (scratch buffers are not on this path)
static ggml_tensor *ggml_new_tensor_impl(ggml_context *ctx, int type, int n_dims, const int64_t *ne, void *data)
{
    ggml_object *obj_cur = ctx->objects_end;
    const uint64_t cur_offs = obj_cur ? obj_cur->offs : 0, cur_size = obj_cur ? obj_cur->size : 0;
    const uint64_t cur_end = cur_offs + cur_size;
    uint64_t size_needed = 0;
    if (!data && !ctx->no_alloc) {
        size_needed += TSIZE[type] * (uint64_t)(ne[0] / BLCK[type]);
        for (int i = 1; i < n_dims; i++) size_needed *= (uint64_t)ne[i];
        size_needed = ((size_needed + GGML_MEM_ALIGN - 1) / GGML_MEM_ALIGN) * GGML_MEM_ALIGN;
    }
    size_needed += sizeof(ggml_tensor);
    if (cur_end + size_needed + GGML_OBJECT_SIZE > ctx->mem_size) {
        fprintf(stderr, "ggml_new_tensor_impl: not enough space in the context's memory pool (needed %llu, available %llu)\n",
                (unsigned long long)(cur_end + size_needed + GGML_OBJECT_SIZE), (unsigned long long)ctx->mem_size);
        return nullptr;                                          // Ggml.cs:7757-7763
    }
    uint8_t *mem = static_cast<uint8_t *>(ctx->mem_buffer);
    ggml_object *obj_new = reinterpret_cast<ggml_object *>(mem + cur_end);
    obj_new->offs = cur_end + GGML_OBJECT_SIZE;
    obj_new->size = size_needed;
    obj_new->next = nullptr;
    if (obj_cur) obj_cur->next = obj_new; else ctx->objects_begin = obj_new;
    ctx->objects_end = obj_new;

    ggml_tensor *r = reinterpret_cast<ggml_tensor *>(mem + obj_new->offs);
    memset(r, 0, sizeof *r);
    r->type = type;
    r->n_dims = n_dims;
    r->op = GGML_OP_NONE;
    r->data = (!data && !ctx->no_alloc) ? static_cast<void *>(r + 1) : data;
    for (int i = 0; i < GGML_MAX_DIMS; i++) r->ne[i] = 1;
    for (int i = 0; i < n_dims; i++) r->ne[i] = ne[i];
    r->nb[0] = TSIZE[type];
    r->nb[1] = r->nb[0] * (uint64_t)(r->ne[0] / BLCK[type]);
    for (int i = 2; i < GGML_MAX_DIMS; i++) r->nb[i] = r->nb[i - 1] * (uint64_t)r->ne[i - 1];
    ctx->n_objects++;
    return r;
}

ggml_tensor *ggml_new_tensor(ggml_context *ctx, int type, int n_dims, const int64_t *ne) { return ggml_new_tensor_impl(ctx, type, n_dims, ne, nullptr); }
ggml_tensor *ggml_new_tensor_1d(ggml_context *ctx, int type, int64_t ne0) { return ggml_new_tensor(ctx, type, 1, &ne0); }
ggml_tensor *ggml_new_tensor_2d(ggml_context *ctx, int type, int64_t ne0, int64_t ne1) { const int64_t ne[2] = {ne0, ne1}; return ggml_new_tensor(ctx, type, 2, ne); }
ggml_tensor *ggml_new_tensor_3d(ggml_context *ctx, int type, int64_t ne0, int64_t ne1, int64_t ne2) { const int64_t ne[3] = {ne0, ne1, ne2}; return ggml_new_tensor(ctx, type, 3, ne); }
ggml_tensor *ggml_new_tensor_4d(ggml_context *ctx, int type, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3) { const int64_t ne[4] = {ne0, ne1, ne2, ne3}; return ggml_new_tensor(ctx, type, 4, ne); }

ggml_tensor *ggml_set_f32(ggml_tensor *t, float value)
{
    if (t->type != GGML_TYPE_F32) { g_last_status = GGB_E_UNSUPPORTED; return t; }
    const int64_t nr = ggml_nrows(t);
    for (int64_t r = 0; r < nr; r++) { float *p = reinterpret_cast<float *>(static_cast<char *>(t->data) + r * t->nb[1]); for (int64_t i = 0; i < t->ne[0]; i++) p[i] = value; }
    // a setter of the API rewrites tensor->data: a resident device copy of these bytes (weight cache, opt-in) is stale now
    if (ggb_pool *pool = pool_holding(t->data)) ggb_tensor_invalidate(pool, t);
    return t;
}
float ggml_get_f32_1d(const ggml_tensor *t, int i) { return t->type == GGML_TYPE_F32 ? static_cast<const float *>(t->data)[i] : 0.0f; }

static bool ggml_can_mul_mat(const ggml_tensor *a, const ggml_tensor *b) { return a->ne[0] == b->ne[0] && a->ne[2] == b->ne[2] && a->ne[3] == b->ne[3]; }

// Ggml.cs:8222-8246.  (The autodiff branch -- is_node / grad -- is outside this path.)
ggml_tensor *ggml_mul_mat(ggml_context *ctx, ggml_tensor *a, ggml_tensor *b)
{
    if (!ggml_can_mul_mat(a, b) || a->nb[0] > a->nb[1]) {        // Debug.Assert in the reference
        fprintf(stderr, "ggml_mul_mat: ggml_can_mul_mat / !ggml_is_transposed failed\n");
        g_last_status = GGB_E_INVALID;
        return nullptr;
    }
    const int64_t ne[4] = {a->ne[1], b->ne[1], a->ne[2], b->ne[3]};
    ggml_tensor *r = ggml_new_tensor(ctx, GGML_TYPE_F32, a->n_dims < b->n_dims ? a->n_dims : b->n_dims, ne);
    if (!r) return nullptr;
    r->op = GGML_OP_MUL_MAT;
    r->src0 = a;
    r->src1 = b;
    return r;
}

// Ggml.cs:8275-8305: the result is a view of b with op = CPY, src0 = a, src1 = b.
ggml_tensor *ggml_cpy(ggml_context *ctx, ggml_tensor *a, ggml_tensor *b)
{
    if (ggml_nelements(a) != ggml_nelements(b)) { g_last_status = GGB_E_INVALID; return nullptr; }
    ggml_tensor *r = ggml_new_tensor_impl(ctx, b->type, b->n_dims, b->ne, b->data);     // ggml_view_tensor
    if (!r) return nullptr;
    for (int i = 0; i < GGML_MAX_DIMS; i++) r->nb[i] = b->nb[i];
    r->op = GGML_OP_CPY;
    r->src0 = a;
    r->src1 = b;
    return r;
}

// ---- the neighbours of mul_mat in a Llama layer: builders only, the arithmetic is in libggb200.so ----

static bool ggml_are_same_shape(const ggml_tensor *a, const ggml_tensor *b) { return a->ne[0] == b->ne[0] && a->ne[1] == b->ne[1] && a->ne[2] == b->ne[2] && a->ne[3] == b->ne[3]; }
// ggml_dup_tensor / ggml_view_tensor (Ggml.cs:2397-2406, 2850-2861)
ggml_tensor *ggml_dup_tensor(ggml_context *ctx, const ggml_tensor *src) { return ggml_new_tensor_impl(ctx, src->type, src->n_dims, src->ne, nullptr); }
ggml_tensor *ggml_view_tensor(ggml_context *ctx, const ggml_tensor *src)
{
    ggml_tensor *r = ggml_new_tensor_impl(ctx, src->type, src->n_dims, src->ne, src->data);
    if (r) for (int i = 0; i < GGML_MAX_DIMS; i++) r->nb[i] = src->nb[i];
    return r;
}
static ggml_tensor *unary_or_binary(ggml_context *ctx, int op, ggml_tensor *a, ggml_tensor *b, bool inplace)
{
    ggml_tensor *r = inplace ? ggml_view_tensor(ctx, a) : ggml_dup_tensor(ctx, a);
    if (!r) return nullptr;
    r->op = op; r->src0 = a; r->src1 = b;
    return r;
}
// Ggml.cs:7868-7890, 7918-7945: Debug.Assert(ggml_are_same_shape(a, b))
ggml_tensor *ggml_add(ggml_context *ctx, ggml_tensor *a, ggml_tensor *b)
{
    if (!ggml_are_same_shape(a, b)) { g_last_status = GGB_E_INVALID; return nullptr; }
    return unary_or_binary(ctx, GGML_OP_ADD, a, b, false);
}
ggml_tensor *ggml_add_inplace(ggml_context *ctx, ggml_tensor *a, ggml_tensor *b)
{
    if (!ggml_are_same_shape(a, b)) { g_last_status = GGB_E_INVALID; return nullptr; }
    return unary_or_binary(ctx, GGML_OP_ADD, a, b, true);
}
ggml_tensor *ggml_mul(ggml_context *ctx, ggml_tensor *a, ggml_tensor *b)
{
    if (!ggml_are_same_shape(a, b)) { g_last_status = GGB_E_INVALID; return nullptr; }
    return unary_or_binary(ctx, GGML_OP_MUL, a, b, false);
}
// Ggml.cs:6911-6922, 7978-8000 -- NOT on the B200 path: a node the caller's own CPU loop runs (see ggml_graph_compute below)
ggml_tensor *ggml_sqr(ggml_context *ctx, ggml_tensor *a) { return unary_or_binary(ctx, GGML_OP_SQR_, a, nullptr, false); }
ggml_tensor *ggml_sqr_inplace(ggml_context *ctx, ggml_tensor *a) { return unary_or_binary(ctx, GGML_OP_SQR_, a, nullptr, true); }
ggml_tensor *ggml_silu(ggml_context *ctx, ggml_tensor *a) { return unary_or_binary(ctx, GGML_OP_SILU, a, nullptr, false); }                 // Ggml.cs:8154-8174
ggml_tensor *ggml_silu_inplace(ggml_context *ctx, ggml_tensor *a) { return unary_or_binary(ctx, GGML_OP_SILU, a, nullptr, true); }
ggml_tensor *ggml_rms_norm(ggml_context *ctx, ggml_tensor *a) { return unary_or_binary(ctx, GGML_OP_RMS_NORM, a, nullptr, false); }         // Ggml.cs:8199-8220
ggml_tensor *ggml_cont(ggml_context *ctx, ggml_tensor *a) { return unary_or_binary(ctx, GGML_OP_CONT, a, nullptr, false); }                 // Ggml.cs:8301-8322
// Ggml.cs:8248-8271: the result is ALWAYS a view of a (the reference's TODO), so the node scales a's data in place
ggml_tensor *ggml_scale(ggml_context *ctx, ggml_tensor *a, ggml_tensor *b)
{
    if (ggml_nelements(b) != 1) { g_last_status = GGB_E_INVALID; return nullptr; }
    return unary_or_binary(ctx, GGML_OP_SCALE, a, b, true);
}
// Ggml.cs:8015-8040: result has b's shape; ggml_can_repeat (8398-8407)
ggml_tensor *ggml_repeat(ggml_context *ctx, ggml_tensor *a, ggml_tensor *b)
{
    for (int i = 0; i < GGML_MAX_DIMS; i++) if (a->ne[i] <= 0 || b->ne[i] % a->ne[i]) { g_last_status = GGB_E_INVALID; return nullptr; }
    if (ggml_are_same_shape(a, b)) return a;
    ggml_tensor *r = ggml_new_tensor(ctx, a->type, b->n_dims, b->ne);
    if (!r) return nullptr;
    r->op = GGML_OP_REPEAT; r->src0 = a; r->src1 = b;
    return r;
}
// Ggml.cs:7199-7225: a view with ne0/ne1 and nb0/nb1 swapped
ggml_tensor *ggml_transpose(ggml_context *ctx, ggml_tensor *a)
{
    ggml_tensor *r = ggml_view_tensor(ctx, a);
    if (!r) return nullptr;
    r->ne[0] = a->ne[1]; r->ne[1] = a->ne[0];
    r->nb[0] = a->nb[1]; r->nb[1] = a->nb[0];
    r->op = GGML_OP_TRANSPOSE; r->src0 = a; r->src1 = nullptr;
    return r;
}

// Ggml.cs:7559-7623
static void ggml_visit_parents(ggml_cgraph *g, ggml_tensor *node)
{
    for (int i = 0; i < g->n_nodes; i++) if (g->nodes[i] == node) return;
    for (int i = 0; i < g->n_leafs; i++) if (g->leafs[i] == node) return;
    if (node->src0) ggml_visit_parents(g, node->src0);
    if (node->src1) ggml_visit_parents(g, node->src1);
    for (int i = 0; i < GGML_MAX_OPT; i++) if (node->opt[i]) ggml_visit_parents(g, reinterpret_cast<ggml_tensor *>(node->opt[i]));
    if (node->op == GGML_OP_NONE && !node->grad) {
        if (g->n_leafs < GGML_MAX_NODES) g->leafs[g->n_leafs++] = node;
    } else if (g->n_nodes < GGML_MAX_NODES) {
        g->nodes[g->n_nodes] = node;
        g->grads[g->n_nodes] = node->grad;
        g->n_nodes++;
    }
}

void ggml_build_forward_expand(ggml_cgraph *g, ggml_tensor *tensor) { ggml_visit_parents(g, tensor); }

// The reference returns the 98 KB struct by value (Ggml.cs:7653); C callers pass the destination.
void ggml_build_forward_into(ggml_cgraph *g, ggml_tensor *tensor)
{
    memset(g, 0, sizeof *g);
    g->n_threads = GGML_DEFAULT_N_THREADS;
    ggml_visit_parents(g, tensor);
}
ggml_cgraph ggml_build_forward(ggml_tensor *tensor) { ggml_cgraph g; ggml_build_forward_into(&g, tensor); return g; }

// Ggml.cs:3209-3736.  The planner's work-buffer sizing is kept (the reference allocates the buffer as an I8
// tensor in the user's context, Ggml.cs:3526-3533, so arena consumption stays identical); the per-node
// INIT/COMPUTE/FINALIZE thread-pool loop is replaced by one call into the CUDA-stream executor.
void ggml_graph_compute(ggml_context *ctx, ggml_cgraph *g)
{
    const int n_threads = g->n_threads;
    uint64_t work_size = 0;
    for (int i = 0; i < g->n_nodes; i++) {
        ggml_tensor *node = g->nodes[i];
        if (node->op == GGML_OP_MUL_MAT) {
            node->n_tasks = n_threads;                           // Ggml.cs:3331
            uint64_t cur = 0;
            const int t0 = node->src0->type;
            if (t0 == GGML_TYPE_F16) cur = TSIZE[GGML_TYPE_F16] * (uint64_t)ggml_nelements(node->src1);
            // vec_dot_type of quantize_fns[] looked up by type (defect D1 repaired): Ggml.cs:226, 235, 244, 254, 263, 272
            else if (t0 == GGML_TYPE_Q4_0 || t0 == GGML_TYPE_Q4_2 || t0 == GGML_TYPE_Q5_0 || t0 == GGML_TYPE_Q8_0)
                cur = TSIZE[GGML_TYPE_Q8_0] * (uint64_t)ggml_nelements(node->src1) / 32;
            else if (t0 == GGML_TYPE_Q4_1 || t0 == GGML_TYPE_Q5_1) cur = TSIZE[GGML_TYPE_Q8_1] * (uint64_t)ggml_nelements(node->src1) / 32;
            if (cur > work_size) work_size = cur;
        } else if (node->op == GGML_OP_CPY) {
            node->n_tasks = n_threads;                           // Ggml.cs:3266-3286
        } else {
            node->n_tasks = 1;
        }
    }
    if (work_size > 0 && !g->work) {
        g->work_size = work_size + (uint64_t)CACHE_LINE_SIZE * (uint64_t)(n_threads - 1);
        g->work = ggml_new_tensor_1d(ctx, GGML_TYPE_I8, (int64_t)g->work_size);
    }
    ggb_pool *pool = pool_of(ctx);
    if (!pool) { g_last_status = GGB_E_INVALID; fprintf(stderr, "ggml_graph_compute: unknown context\n"); return; }
    static thread_local uint8_t done[GGML_MAX_NODES];
    const int rc = ggb_graph_compute_mul_mats(pool, g, 0, done);
    if (rc < 0) { g_last_status = rc; fprintf(stderr, "ggml_graph_compute: %s\n", ggb_last_error()); return; }
    g_last_status = 0;
    for (int i = 0; i < g->n_nodes; i++)
        if (!done[i]) {
            // RESHAPE/VIEW/PERMUTE/TRANSPOSE are no-ops in the reference (Ggml.cs:8668-8687); anything else is off-path here
            ggml_tensor *node = g->nodes[i];
            const int op = node->op;
            if (op >= 24 && op <= 27) continue;
            // What the unchanged C# loop does with the nodes the backend left alone (Ggml.cs:3539-3704), in node order.  Two ops are
            // implemented here so that tests can interleave host-side nodes with device nodes: SQR (Ggml.cs:5109-5130) and the
            // same-type contiguous CPY (a memcpy, Ggml.cs:4208-4213).  This is the HOST's work, not a fallback of the device path.
            if (op == GGML_OP_SQR_ && node->type == GGML_TYPE_F32 && node->src0 && node->src0->type == GGML_TYPE_F32) {
                const float *x = static_cast<const float *>(node->src0->data); float *y = static_cast<float *>(node->data);
                const int64_t ne = ggml_nelements(node);
                for (int64_t k = 0; k < ne; k++) y[k] = x[k] * x[k];
                continue;
            }
            if (op == GGML_OP_CPY && node->src0 && node->src1 && node->src0->type == node->src1->type && ggml_nbytes(node->src0) == ggml_nbytes(node->src1)) {
                memmove(node->src1->data, node->src0->data, ggml_nbytes(node->src0));
                continue;
            }
            g_last_status = GGB_E_UNSUPPORTED;
            fprintf(stderr, "ggml_graph_compute: node %d (op %d) is outside the B200 mul_mat path; the C# CPU loop would run it\n", i, op);
        }
}

} // extern "C"
