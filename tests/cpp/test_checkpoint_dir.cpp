// CPU test of the checkpoint directory protocol (incubator_pegasus_b200/host/checkpoint_dir.h).  Built and run by tests/test_checkpoint_dir.py.
#include <cstdlib>

#include "../../incubator_pegasus_b200/host/checkpoint_dir.h"

using namespace pgs;
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "%s:%d: CHECK(%s) failed\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

static bool exists(const std::string &p) { struct stat sb; return stat(p.c_str(), &sb) == 0; }

int main(int argc, char **argv)
{
    CHECK(argc == 2);
    const std::string base = std::string(argv[1]) + "/replica";
    CheckpointManifest m;
    m.app_id = 3; m.pidx = 7; m.data_version = 1; m.decree = 42; m.last_seq = 1234;
    m.files = {{2, "000001.sst", 5}, {0, "000002.sst", 3}};
    {   // an attempt that dies before the manifest: nothing of the final name exists, the next attempt starts clean
        CheckpointWriter w;
        CHECK(w.begin(base, 42) == 0);
        CHECK(w.add_file("000001.sst", "hello", 5));
        CHECK(exists(base + "/checkpoint.42.tmp/000001.sst") && !exists(base + "/checkpoint.42"));
        CHECK(!w.add_file("../evil", "x", 1) && !w.add_file("", "x", 1));
    }   // destructor = abandon
    CHECK(!exists(base + "/checkpoint.42.tmp") && !exists(base + "/checkpoint.42"));
    {   // leftovers of a crashed process (no destructor ran) and of the pre-rename layout are cleared
        CHECK(mkdir((base + "/checkpoint.42.tmp").c_str(), 0755) == 0 && ckpt_write_file(base + "/checkpoint.42.tmp/junk", "j", 1));
        CHECK(mkdir((base + "/checkpoint.42").c_str(), 0755) == 0 && ckpt_write_file(base + "/checkpoint.42/000001.sst", "half", 4));
        CHECK(!ckpt_is_complete(base + "/checkpoint.42"));
        CheckpointWriter w;
        CHECK(w.begin(base, 42) == 0);
        CHECK(!exists(base + "/checkpoint.42") && !exists(base + "/checkpoint.42.tmp/junk"));
        CHECK(w.add_file("000001.sst", "hello", 5) && w.add_file("000002.sst", "abc", 3));
        CHECK(w.commit(m.str()));
        CHECK(w.dir() == base + "/checkpoint.42");
    }
    CHECK(ckpt_is_complete(base + "/checkpoint.42") && !exists(base + "/checkpoint.42.tmp"));
    {   // the same decree again: already there
        CheckpointWriter w;
        CHECK(w.begin(base, 42) == 1);
    }
    std::vector<uint8_t> raw;
    CHECK(ckpt_read_file(base + "/checkpoint.42/MANIFEST", raw));
    CheckpointManifest r;
    CHECK(r.parse(std::string(raw.begin(), raw.end())));
    CHECK(r.app_id == 3 && r.pidx == 7 && r.data_version == 1 && r.decree == 42 && r.last_seq == 1234 && r.files.size() == 2);
    CHECK(r.files[0].level == 2 && r.files[0].name == "000001.sst" && r.files[0].bytes == 5 && r.files[1].level == 0);
    CHECK(ckpt_read_file(base + "/checkpoint.42/000002.sst", raw) && std::string(raw.begin(), raw.end()) == "abc");
    // manifests that must be refused
    const std::string good = m.str();
    CHECK(r.parse(good));
    for (const std::string &bad : {std::string(""), std::string("pegasus_b200_checkpoint 2\n"), good.substr(0, good.size() - 8),
                                   std::string("pegasus_b200_checkpoint 1\napp_id 1\npidx 0\ndata_version 1\nlast_flushed_decree 5\nlast_seq 1\nruns 99999999999\n"),
                                   std::string("pegasus_b200_checkpoint 1\napp_id 1\npidx 0\ndata_version 1\nlast_flushed_decree 5\nlast_seq 1\nruns 1\n0 ../x 3\n"),
                                   std::string("pegasus_b200_checkpoint 1\napp_id 1\npidx 0\ndata_version 1\nlast_flushed_decree -5\nlast_seq 1\nruns 0\n"),
                                   std::string("pegasus_b200_checkpoint 1\napp_id 1\npidx 0\nlast_seq 1\ndata_version 1\nlast_flushed_decree 5\nruns 0\n"),
                                   std::string("pegasus_b200_checkpoint 1\napp_id 1\npidx 0\ndata_version 1\nlast_flushed_decree 5\nlast_seq 1\nruns 1\n-1 a.sst 3\n")})
        CHECK(!r.parse(bad));
    CHECK(ckpt_remove_flat_dir(base + "/checkpoint.42") && ckpt_remove_flat_dir(base + "/nothing-here"));
    printf("OK\n");
    return 0;
}
