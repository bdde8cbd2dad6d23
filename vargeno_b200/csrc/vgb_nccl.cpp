// vgb_nccl.cpp -- the one exchange step of the path: sum of the per-GPU pileup counters.
//
// NCCL is reached through dlopen so that libvgb200.so has no link-time NCCL dependency: inside a PyTorch process
// the already loaded libnccl.so.2 is reused, in the stand-alone C++ host the system one is opened.
#include <dlfcn.h>

#include <cstring>

#include "vgb_internal.h"

namespace {

struct NcclUniqueId { char internal[128]; };      // ncclUniqueId (nccl.h)
typedef void *NcclComm;
typedef int (*fn_get_uid)(NcclUniqueId *);
typedef int (*fn_init_rank)(NcclComm *, int, NcclUniqueId, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_destroy)(NcclComm);
typedef const char *(*fn_errstr)(int);

struct Api {
	void *h = nullptr;
	fn_get_uid get_uid = nullptr;
	fn_init_rank init_rank = nullptr;
	fn_allreduce allreduce = nullptr;
	fn_destroy destroy = nullptr;
	fn_errstr errstr = nullptr;
};

Api *api(std::string &err)
{
	static Api a;
	if (a.h) return &a;
	const char *names[] = { "libnccl.so.2", "libnccl.so", nullptr };
	for (int i = 0; names[i] && !a.h; i++) a.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
	if (!a.h) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return nullptr; }
	a.get_uid = (fn_get_uid)dlsym(a.h, "ncclGetUniqueId");
	a.init_rank = (fn_init_rank)dlsym(a.h, "ncclCommInitRank");
	a.allreduce = (fn_allreduce)dlsym(a.h, "ncclAllReduce");
	a.destroy = (fn_destroy)dlsym(a.h, "ncclCommDestroy");
	a.errstr = (fn_errstr)dlsym(a.h, "ncclGetErrorString");
	if (!a.get_uid || !a.init_rank || !a.allreduce || !a.destroy) { err = "libnccl lacks the expected symbols"; a.h = nullptr; return nullptr; }
	return &a;
}

const int NCCL_UINT32 = 3;   // ncclUint32
const int NCCL_SUM = 0;      // ncclSum

}  // namespace

namespace vgb {

int nccl_unique_id(void *out128, std::string &err)
{
	Api *a = api(err);
	if (!a) return VGB_E_NCCL;
	NcclUniqueId id;
	const int r = a->get_uid(&id);
	if (r) { err = std::string("ncclGetUniqueId: ") + (a->errstr ? a->errstr(r) : "error"); return VGB_E_NCCL; }
	memcpy(out128, &id, 128);
	return VGB_OK;
}

int nccl_init(vgb_ctx *c)
{
	std::string err;
	Api *a = api(err);
	if (!a) return set_err(c, VGB_E_NCCL, "%s", err.c_str());
	NcclUniqueId id;
	memcpy(&id, c->cfg.nccl_unique_id, 128);
	NcclComm comm = nullptr;
	const int r = a->init_rank(&comm, c->cfg.world_size, id, c->cfg.rank);
	if (r) return set_err(c, VGB_E_NCCL, "ncclCommInitRank: %s", a->errstr ? a->errstr(r) : "error");
	c->nccl_comm = comm;
	return VGB_OK;
}

int nccl_allreduce_u32(vgb_ctx *c, uint32_t *buf, uint64_t n)
{
	std::string err;
	Api *a = api(err);
	if (!a || !c->nccl_comm) return set_err(c, VGB_E_NCCL, "NCCL communicator not initialised");
	const int r = a->allreduce(buf, buf, (size_t)n, NCCL_UINT32, NCCL_SUM, c->nccl_comm, c->stream);
	if (r) return set_err(c, VGB_E_NCCL, "ncclAllReduce: %s", a->errstr ? a->errstr(r) : "error");
	return VGB_OK;
}

void nccl_destroy(vgb_ctx *c)
{
	std::string err;
	Api *a = api(err);
	if (a && c->nccl_comm) a->destroy(c->nccl_comm);
	c->nccl_comm = nullptr;
}

}  // namespace vgb
