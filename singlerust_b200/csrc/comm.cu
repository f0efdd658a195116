// comm.cu — C1: the one exchange step of the cell-row-sharded job: sum-allreduce of the per-gene integer
// moment limbs (1.4 MB) and of the d x d Gram matrix (32 MB at d = 2048) over NCCL / NVLink. The reference is
// single-process; this is new. NCCL is resolved at run time (dlopen "libnccl.so.2"), so the library loads on a
// box without NCCL and in a process where torch already mapped its own copy.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace srb {

typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
enum { NCCL_SUM = 0, NCCL_PROD = 1, NCCL_MAX = 2, NCCL_MIN = 3 };
enum { NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 };  // ncclDataType_t values (nccl.h): ncclUint64 = 5, ncclFloat64 = 8

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi &nccl() {
    static NcclApi api;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!api.lib) {
        api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) throw Error(SRB_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
        api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.AllGather)
            throw Error(SRB_ERR_NCCL, "libnccl.so.2 lacks a required symbol");
    }
    return api;
}

#define SRB_NCCL(expr)                                                                                       \
    do {                                                                                                     \
        int _e = (expr);                                                                                     \
        if (_e != 0) {                                                                                       \
            const char *es = nccl().GetErrorString ? nccl().GetErrorString(_e) : "?";                        \
            throw srb::Error(SRB_ERR_NCCL, std::string(#expr) + ": " + es);                                  \
        }                                                                                                    \
    } while (0)

static void allreduce(srb_ctx *ctx, void *buf, size_t n, int dtype, int op) {
    if (ctx->nranks <= 1 || n == 0) return;
    SRB_REQUIRE(ctx->comm, SRB_ERR_NCCL, "communicator not initialised");
    SRB_NCCL(nccl().AllReduce(buf, buf, n, dtype, op, (nccl_comm)ctx->comm, ctx->stream));
}
void allreduce_u64_sum(srb_ctx *ctx, uint64_t *d, size_t n) { allreduce(ctx, d, n, NCCL_UINT64, NCCL_SUM); }
void allreduce_f64_sum(srb_ctx *ctx, double *d, size_t n) { allreduce(ctx, d, n, NCCL_FLOAT64, NCCL_SUM); }
void allreduce_f64_min(srb_ctx *ctx, double *d, size_t n) { allreduce(ctx, d, n, NCCL_FLOAT64, NCCL_MIN); }
void allreduce_f64_max(srb_ctx *ctx, double *d, size_t n) { allreduce(ctx, d, n, NCCL_FLOAT64, NCCL_MAX); }

// in place: rank r's `count` doubles sit at buf + r * count on entry; on exit every rank holds all nranks * count
void allgather_f64(srb_ctx *ctx, cudaStream_t stream, double *buf, size_t count) {
    if (ctx->nranks <= 1 || count == 0) return;
    SRB_REQUIRE(ctx->comm, SRB_ERR_NCCL, "communicator not initialised");
    SRB_NCCL(nccl().AllGather(buf + (size_t)ctx->rank * count, buf, count, NCCL_FLOAT64, (nccl_comm)ctx->comm, stream));
}

void comm_destroy(srb_ctx *ctx) {
    if (ctx->comm) nccl().CommDestroy((nccl_comm)ctx->comm);
    ctx->comm = nullptr;
    ctx->nranks = 1, ctx->rank = 0;
}

}  // namespace srb

using namespace srb;

extern "C" {

int32_t srb_comm_unique_id(void *id128) {
    SRB_API_BEGIN
    SRB_REQUIRE(id128, SRB_ERR_INVALID_ARG, "id128 is null");
    SRB_NCCL(nccl().GetUniqueId((nccl_uid *)id128));
    SRB_API_END
}

int32_t srb_ctx_comm_init(srb_ctx *ctx, const void *id128, int32_t rank, int32_t nranks) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx && id128, SRB_ERR_INVALID_ARG, "null argument");
    SRB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, SRB_ERR_INVALID_ARG, "bad rank / nranks");
    SRB_CUDA(cudaSetDevice(ctx->device));
    comm_destroy(ctx);
    if (nranks > 1) {
        nccl_uid id;
        memcpy(&id, id128, sizeof(id));
        nccl_comm comm = nullptr;
        SRB_NCCL(nccl().CommInitRank(&comm, nranks, id, rank));
        ctx->comm = comm;
    }
    ctx->rank = rank, ctx->nranks = nranks;
    SRB_API_END
}

}  // extern "C"
