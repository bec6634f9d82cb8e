// Layer-plan executor and the library-level C ABI (errors, launch counter, device info).
//
// A plan is the flat op list the Python host derives from a model description (celldetection_b200/models/graph.py):
// it replaces the nn.Module call tree of CPNCore.forward (/root/reference/celldetection/models/cpn.py:238-283) with
// one native loop over pre-validated kernel launches on a caller-provided arena; no device allocation, no host
// synchronisation, TMA tensor maps built once at plan creation.
#include "common.cuh"
#include <string>
#include <vector>
#include <cstring>

namespace cpn {

static thread_local std::string g_error = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}

int sm_count() {
  static int cached = 0;
  if (cached <= 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
    cached = prop.multiProcessorCount;
  }
  return cached;
}

}  // namespace cpn

using namespace cpn;

struct cpn_plan {
  std::vector<cpn_op_t> ops;
  std::vector<ConvTcPlan*> tc;  // per op (nullptr unless TCGEN05 conv)
  const char* weights;
  size_t weights_bytes;
  char* arena;
  size_t arena_bytes;
  int32_t* flags;
  int n_launches;
};

extern "C" int cpn_abi_version(void) { return CPN_B200_ABI_VERSION; }
extern "C" const char* cpn_last_error(void) { return g_error.c_str(); }
extern "C" int64_t cpn_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int cpn_device_info(char* name_host, int n, int* sm_count_host, int* cc_major_host, int* cc_minor_host) {
  int dev = 0;
  CPN_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CPN_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (name_host && n > 0) {
    strncpy(name_host, prop.name, (size_t)n - 1);
    name_host[n - 1] = 0;
  }
  if (sm_count_host) *sm_count_host = prop.multiProcessorCount;
  if (cc_major_host) *cc_major_host = prop.major;
  if (cc_minor_host) *cc_minor_host = prop.minor;
  return 0;
}

static size_t view_bytes(const cpn_view_t& v) {
  if (v.n == 0) return 0;
  const size_t last = (size_t)v.c + (dtype_has_lo(v.dtype) ? (size_t)v.lo_delta : 0);   // one past the last element
  return ((size_t)v.n * v.h * v.w - 1) * (size_t)v.pitch * dtype_size(v.dtype) + last * dtype_size(v.dtype);
}

static int check_view(const cpn_view_t& v, size_t limit, const char* what, int i) {
  CPN_REQUIRE(v.n > 0 && v.h > 0 && v.w > 0 && v.c > 0 && v.pitch >= v.c, "op %d: bad %s view (%d,%d,%d,%d pitch %d)", i,
              what, v.n, v.h, v.w, v.c, v.pitch);
  CPN_REQUIRE(!dtype_has_lo(v.dtype) || (v.lo_delta >= v.c && v.lo_delta + v.c <= v.pitch),
              "op %d: bad %s split view (c %d lo_delta %d pitch %d)", i, what, v.c, v.lo_delta, v.pitch);
  CPN_REQUIRE(v.dtype != CPN_DT_F16F8 || (v.lo_delta % 32 == 0 && v.c % 32 == 0 && v.fp8_exp >= -64 && v.fp8_exp <= 64),
              "op %d: bad %s fp16+e4m3 view (c %d lo_delta %d fp8_exp %d)", i, what, v.c, v.lo_delta, v.fp8_exp);
  CPN_REQUIRE(v.offset >= 0 && (size_t)v.offset + view_bytes(v) <= limit,
              "op %d: %s view [%lld, +%zu) exceeds its buffer of %zu bytes", i, what, (long long)v.offset,
              view_bytes(v), limit);
  return 0;
}

extern "C" int cpn_plan_create(const cpn_op_t* ops_host, int n_ops, const void* weights, size_t weights_bytes,
                               void* arena, size_t arena_bytes, int32_t* flags_dev, cpn_plan_t** plan_out) {
  CPN_REQUIRE(ops_host && n_ops > 0 && plan_out, "plan_create: bad arguments");
  CPN_REQUIRE(arena && flags_dev, "plan_create: arena and flags must be device pointers");
  cpn_plan* pl = new cpn_plan();
  pl->ops.assign(ops_host, ops_host + n_ops);
  pl->tc.assign(n_ops, nullptr);
  pl->weights = reinterpret_cast<const char*>(weights);
  pl->weights_bytes = weights_bytes;
  pl->arena = reinterpret_cast<char*>(arena);
  pl->arena_bytes = arena_bytes;
  pl->flags = flags_dev;
  pl->n_launches = 0;
  auto fail = [&]() { cpn_plan_destroy(pl); return 1; };
  for (int i = 0; i < n_ops; ++i) {
    const cpn_op_t& op = pl->ops[i];
    const bool dst_bound = op.out_binding >= 0;
    if (!dst_bound && check_view(op.dst, arena_bytes, "dst", i)) return fail();
    if (op.kind != CPN_OP_PREP && check_view(op.src, arena_bytes, "src", i)) return fail();
    if (op.kind == CPN_OP_CONV || op.kind == CPN_OP_PROJ) {
      if (!(op.w_offset >= 0 && (size_t)op.w_offset < weights_bytes && op.b_offset < (int64_t)weights_bytes)) {
        set_error("op %d: weight offsets out of range", i);
        return fail();
      }
    }
    if (op.kind == CPN_OP_CONV) {
      if (op.res.n && check_view(op.res, arena_bytes, "res", i)) return fail();
      if (op.engine == CPN_ENGINE_TCGEN05) {
        if (dst_bound) { set_error("op %d: tcgen05 conv cannot write a bound output", i); return fail(); }
        const float* bias = op.b_offset >= 0 ? reinterpret_cast<const float*>(pl->weights + op.b_offset) : nullptr;
        if (conv_tc_plan_create(op, pl->arena + op.src.offset, pl->arena + op.dst.offset,
                                op.res.n ? pl->arena + op.res.offset : nullptr, pl->weights + op.w_offset, bias,
                                &pl->tc[i]))
          return fail();
        if (op.fuse_next > 0) {   // the next `fuse_next` PROJ ops run inside this convolution's epilogue
          if (i + op.fuse_next > n_ops - 1) { set_error("op %d: fuse_next out of range", i); return fail(); }
          if (conv_tc_fuse_proj(pl->tc[i], op.fuse_next, &pl->ops[i + 1], pl->weights)) return fail();
        }
      } else if (op.engine != CPN_ENGINE_SIMT) {
        set_error("op %d: unknown conv engine %d", i, op.engine);
        return fail();
      }
    }
    if (op.kind < CPN_OP_PREP || op.kind > CPN_OP_PROJ) { set_error("op %d: unknown kind %d", i, op.kind); return fail(); }
    pl->n_launches += 1;
    if (op.kind == CPN_OP_CONV && op.engine == CPN_ENGINE_TCGEN05 && op.fuse_next > 0) pl->n_launches -= op.fuse_next;
  }
  *plan_out = pl;
  return 0;
}

static int run_op(cpn_plan* pl, int i, const void* input, int input_format, void* const* outputs, int n_outputs,
                  cudaStream_t st) {
  const cpn_op_t& op = pl->ops[i];
  void* dst;
  if (op.out_binding >= 0) {
    CPN_REQUIRE(op.out_binding < n_outputs && outputs && outputs[op.out_binding], "op %d: output binding %d not provided",
                i, op.out_binding);
    dst = reinterpret_cast<char*>(outputs[op.out_binding]) + op.dst.offset;
  } else {
    dst = pl->arena + op.dst.offset;
  }
  const void* src = pl->arena + op.src.offset;
  const float* bias = op.b_offset >= 0 ? reinterpret_cast<const float*>(pl->weights + op.b_offset) : nullptr;
  switch (op.kind) {
    case CPN_OP_PREP:
      CPN_REQUIRE(input != nullptr, "forward: input pointer is NULL");
      return prep_launch(op, input, input_format, dst, pl->flags, st);
    case CPN_OP_CONV:
      if (op.engine == CPN_ENGINE_TCGEN05) {
        for (int h = 0; h < op.fuse_next; ++h) {   // bind the fused projections' caller-provided output buffers
          const cpn_op_t& q = pl->ops[i + 1 + h];
          CPN_REQUIRE(q.out_binding < 0 || (q.out_binding < n_outputs && outputs && outputs[q.out_binding]),
                      "op %d: output binding %d not provided", i + 1 + h, q.out_binding);
          char* base = q.out_binding >= 0 ? reinterpret_cast<char*>(outputs[q.out_binding]) : pl->arena;
          conv_tc_bind_proj_out(pl->tc[i], h, base + q.dst.offset);
        }
        return conv_tc_launch(pl->tc[i], st);
      }
      return conv_simt_launch(op, src, dst, op.res.n ? pl->arena + op.res.offset : nullptr, pl->weights + op.w_offset,
                              bias, st);
    case CPN_OP_MAXPOOL: return maxpool_launch(op, src, dst, st);
    case CPN_OP_UPSAMPLE: return upsample_launch(op, src, dst, st);
    case CPN_OP_BILINEAR: return bilinear_launch(op, src, dst, st);
    case CPN_OP_PROJ:
      return proj_launch(op, src, dst, reinterpret_cast<const float*>(pl->weights + op.w_offset), bias, st);
  }
  set_error("op %d: unknown kind", i);
  return 1;
}

extern "C" int cpn_plan_forward(cpn_plan_t* plan, const void* input, int input_format, void* const* outputs_host,
                                int n_outputs, void* stream) {
  CPN_REQUIRE(plan, "plan_forward: NULL plan");
  return cpn_plan_forward_range(plan, 0, (int)plan->ops.size(), input, input_format, outputs_host, n_outputs, stream);
}

extern "C" int cpn_plan_forward_range(cpn_plan_t* plan, int first_op, int end_op, const void* input, int input_format,
                                      void* const* outputs_host, int n_outputs, void* stream) {
  CPN_REQUIRE(plan, "plan_forward_range: NULL plan");
  CPN_REQUIRE(first_op >= 0 && first_op <= end_op && end_op <= (int)plan->ops.size(), "plan_forward_range: bad range [%d, %d)",
              first_op, end_op);
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = first_op; i < end_op; ++i) {
    if (run_op(plan, i, input, input_format, outputs_host, n_outputs, st)) return 1;
    const cpn_op_t& op = plan->ops[i];
    if (op.kind == CPN_OP_CONV && op.engine == CPN_ENGINE_TCGEN05) i += op.fuse_next;  // ran inside the epilogue
  }
  return 0;
}

extern "C" int cpn_plan_run_op(cpn_plan_t* plan, int index, const void* input, int input_format,
                               void* const* outputs_host, int n_outputs, void* stream) {
  CPN_REQUIRE(plan && index >= 0 && index < (int)plan->ops.size(), "plan_run_op: bad index %d", index);
  return run_op(plan, index, input, input_format, outputs_host, n_outputs, (cudaStream_t)stream);
}

extern "C" int cpn_plan_num_launches(const cpn_plan_t* plan) { return plan ? plan->n_launches : 0; }

extern "C" int cpn_plan_set_active_rows(cpn_plan_t* plan, int64_t rows) {
  CPN_REQUIRE(plan, "plan_set_active_rows: NULL plan");
  for (ConvTcPlan* t : plan->tc)
    if (t && conv_tc_limit_rows(t, (long long)rows)) return 1;
  return 0;
}

extern "C" void cpn_plan_destroy(cpn_plan_t* plan) {
  if (!plan) return;
  for (ConvTcPlan* t : plan->tc)
    if (t) conv_tc_plan_destroy(t);
  delete plan;
}

extern "C" int cpn_conv2d(const cpn_op_t* op_host, const void* src_base, void* dst_base, const void* res_base,
                          const void* weights, void* stream) {
  CPN_REQUIRE(op_host && op_host->kind == CPN_OP_CONV, "conv2d: op must be a CONV");
  const cpn_op_t& op = *op_host;
  const char* w = reinterpret_cast<const char*>(weights);
  const float* bias = op.b_offset >= 0 ? reinterpret_cast<const float*>(w + op.b_offset) : nullptr;
  const void* src = reinterpret_cast<const char*>(src_base) + op.src.offset;
  void* dst = reinterpret_cast<char*>(dst_base) + op.dst.offset;
  const void* res = op.res.n ? reinterpret_cast<const char*>(res_base) + op.res.offset : nullptr;
  if (op.engine == CPN_ENGINE_TCGEN05) {
    ConvTcPlan* t = nullptr;
    if (conv_tc_plan_create(op, src, dst, res, w + op.w_offset, bias, &t)) return 1;
    const int rc = conv_tc_launch(t, (cudaStream_t)stream);
    conv_tc_plan_destroy(t);  // launch parameters were copied at launch time
    return rc;
  }
  return conv_simt_launch(op, src, dst, res, w + op.w_offset, bias, (cudaStream_t)stream);
}
