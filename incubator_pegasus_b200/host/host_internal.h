// host_internal.h — C++ declarations shared by the host-side sources of the product.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <string_view>
#include <vector>

namespace pgs {

const uint64_t *crc64_table();
uint64_t crc64(const uint8_t *p, uint64_t n, uint64_t init);
std::string make_key(std::string_view hk, std::string_view sk);
std::string make_next(std::string k);
uint64_t key_hash(std::string_view key);

// Sorted-run builder: RocksDB BlockBuilder + FlushBlockBySizePolicy behaviour (block closes when
// it reached block_size, or when the next entry would overflow it and it is >90 % full); block
// starts padded to 16 bytes.
class RunBuilder
{
public:
    RunBuilder(uint32_t block_size, uint32_t restart_interval)
        : block_size_(block_size), restart_interval_(restart_interval), restarts_(1, 0)
    {
    }
    int32_t add(std::string_view ukey, uint64_t seq, uint8_t type, std::string_view value);
    void finish();
    const std::string &data() const { return data_; }
    const std::vector<uint64_t> &blk_off() const { return blk_off_; }
    const std::vector<uint32_t> &blk_size() const { return blk_size_; }
    uint64_t n_records() const { return n_records_; }

private:
    void flush_block();
    uint32_t block_size_, restart_interval_;
    std::string data_, buf_, last_key_, prev_ukey_;
    std::vector<uint64_t> blk_off_;
    std::vector<uint32_t> blk_size_;
    std::vector<uint32_t> restarts_;
    uint32_t counter_ = 0, entries_ = 0;
    uint64_t n_records_ = 0, prev_trailer_ = 0;
    bool have_prev_ = false;
};

int32_t decode_blocks(const uint8_t *data, const uint64_t *blk_off, const uint32_t *blk_size,
                      uint32_t n_blocks,
                      const std::function<void(std::string_view, uint64_t, uint8_t, std::string_view)> &fn);

// JSON of the user_specified_compaction env -> binary ops table (format.h)
int64_t ops_parse(std::string_view json, uint32_t data_version, std::string &out, uint32_t *n_ops_out);

} // namespace pgs
