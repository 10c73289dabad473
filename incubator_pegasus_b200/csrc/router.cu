// router.cu — box-level placement: one engine per visible GPU, replicas pinned by partition index.
//
// Pegasus tables are hash-partitioned (`pidx = crc64(hash_key) % partition_count`, src/client/partition_resolver.cpp:48-51,
// src/base/pegasus_key_schema.h:150-183) and a replica server hosts many independent replicas
// (src/replica/replica_stub.h: one pegasus_server_impl per gpid).  Nothing on the data path crosses partitions, so a box
// with N GPUs places replica (app_id, pidx) on GPU pidx % N and never moves it: no collective, no peer copies.
#include "engine.h"

#include <vector>

struct pgs_router {
    std::vector<pgs_engine *> engines;
};

extern "C" {

int32_t pgs_router_open(const pgs_engine_config *cfg, int32_t n_devices, pgs_router **out)
{
    if (!out || n_devices < 0) return PGS_INVALID_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) { // no CPU fallback
        pgs::set_error("no CUDA device: %s", cudaGetErrorString(ce));
        return PGS_IO_ERROR;
    }
    if (n_devices == 0) n_devices = ndev;
    if (n_devices > ndev) {
        pgs::set_error("router: %d devices requested, %d visible", n_devices, ndev);
        return PGS_INVALID_ARGUMENT;
    }
    auto *r = new pgs_router;
    for (int d = 0; d < n_devices; d++) {
        pgs_engine_config c{};
        if (cfg) c = *cfg;
        c.device = d;
        pgs_engine *e = nullptr;
        const int32_t st = pgs_engine_open(&c, &e);
        if (st != PGS_OK) {
            for (pgs_engine *x : r->engines) pgs_engine_close(x);
            delete r;
            return st;
        }
        r->engines.push_back(e);
    }
    *out = r;
    return PGS_OK;
}

void pgs_router_close(pgs_router *r)
{
    if (!r) return;
    for (pgs_engine *e : r->engines) pgs_engine_close(e);
    delete r;
}

int32_t pgs_router_device_count(const pgs_router *r) { return r ? (int32_t)r->engines.size() : 0; }

int32_t pgs_router_device_for(const pgs_router *r, int32_t app_id, int32_t pidx)
{
    (void)app_id; // every table spreads the same way: consecutive partitions land on consecutive GPUs
    if (!r || r->engines.empty() || pidx < 0) return -1;
    return pidx % (int32_t)r->engines.size();
}

pgs_engine *pgs_router_engine_for(pgs_router *r, int32_t app_id, int32_t pidx)
{
    const int32_t d = pgs_router_device_for(r, app_id, pidx);
    return d < 0 ? nullptr : r->engines[(size_t)d];
}

uint32_t pgs_partition_index(const uint8_t *hash_key, uint32_t hash_key_len, const uint8_t *sort_key, uint32_t sort_key_len, uint32_t partition_count)
{
    if (!partition_count) return 0;
    // pegasus_key_hash: the hash key decides; an empty hash key falls back to the sort key (pegasus_key_schema.h:150-165)
    const uint64_t h = hash_key_len ? pgs_crc64(hash_key, hash_key_len, 0) : pgs_crc64(sort_key, sort_key_len, 0);
    return (uint32_t)(h % partition_count);
}

} // extern "C"
