// checkpoint_dir.h — the on-disk side of a checkpoint (pegasus_server_impl.cpp:1951-2040 sync_checkpoint; :2200-2336
// storage_apply_checkpoint): files are written into `checkpoint.<decree>.tmp`, synced, and the directory is renamed to
// `checkpoint.<decree>` only when the MANIFEST is in place, so a directory of that name is always complete -- an interrupted
// attempt leaves a .tmp directory that the next attempt clears.  Host-only code (POSIX), unit-tested in tests/cpp.
#pragma once
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include <dirent.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace pgs {

inline bool ckpt_write_file(const std::string &path, const void *p, size_t n)
{
    const int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return false;
    const uint8_t *b = (const uint8_t *)p;
    size_t done = 0;
    while (done < n) {
        const ssize_t w = write(fd, b + done, n - done);
        if (w < 0) { if (errno == EINTR) continue; close(fd); return false; }
        done += (size_t)w;
    }
    const bool ok = fsync(fd) == 0;
    return close(fd) == 0 && ok;
}
inline bool ckpt_read_file(const std::string &path, std::vector<uint8_t> &out)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = n >= 0 && fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}
inline bool ckpt_sync_dir(const std::string &dir)
{
    const int fd = open(dir.c_str(), O_RDONLY | O_DIRECTORY);
    if (fd < 0) return false;
    const bool ok = fsync(fd) == 0;
    close(fd);
    return ok;
}
// remove a directory that holds regular files only (what a checkpoint directory is); true when it is gone or was never there
inline bool ckpt_remove_flat_dir(const std::string &dir)
{
    DIR *d = opendir(dir.c_str());
    if (!d) return errno == ENOENT;
    bool ok = true;
    while (struct dirent *e = readdir(d)) {
        if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
        if (unlink((dir + "/" + e->d_name).c_str()) != 0) ok = false;
    }
    closedir(d);
    return rmdir(dir.c_str()) == 0 && ok;
}
inline bool ckpt_is_complete(const std::string &cdir)
{
    struct stat sb;
    return stat((cdir + "/MANIFEST").c_str(), &sb) == 0 && S_ISREG(sb.st_mode);
}

class CheckpointWriter
{
public:
    // 1 = checkpoint.<decree> is already complete (nothing to do), 0 = started, -1 = I/O error (errno-style text in error())
    int begin(const std::string &parent, int64_t decree)
    {
        final_ = parent + "/checkpoint." + std::to_string(decree);
        tmp_ = final_ + ".tmp";
        if (mkdir(parent.c_str(), 0755) != 0 && errno != EEXIST) return fail("cannot create " + parent);
        if (ckpt_is_complete(final_)) return 1;
        if (!ckpt_remove_flat_dir(final_)) return fail("cannot clear the incomplete " + final_); // a directory without MANIFEST: never valid
        if (!ckpt_remove_flat_dir(tmp_)) return fail("cannot clear " + tmp_);
        if (mkdir(tmp_.c_str(), 0755) != 0) return fail("cannot create " + tmp_);
        open_ = true;
        return 0;
    }
    bool add_file(const std::string &name, const void *p, size_t n)
    {
        if (!open_ || name.empty() || name.find('/') != std::string::npos) { fail("bad file name " + name); return false; }
        if (!ckpt_write_file(tmp_ + "/" + name, p, n)) { fail("cannot write " + tmp_ + "/" + name); return false; }
        return true;
    }
    // the MANIFEST goes last; then the directory gets its final name
    bool commit(const std::string &manifest)
    {
        if (!open_) return false;
        if (!ckpt_write_file(tmp_ + "/MANIFEST", manifest.data(), manifest.size()) || !ckpt_sync_dir(tmp_)) { fail("cannot write the manifest"); return false; }
        if (rename(tmp_.c_str(), final_.c_str()) != 0) { fail("cannot rename " + tmp_); return false; }
        open_ = false;
        const size_t slash = final_.rfind('/');
        ckpt_sync_dir(slash == std::string::npos ? "." : final_.substr(0, slash));
        return true;
    }
    void abandon()
    {
        if (open_) ckpt_remove_flat_dir(tmp_);
        open_ = false;
    }
    ~CheckpointWriter() { abandon(); }
    const std::string &dir() const { return final_; }
    const std::string &error() const { return err_; }

private:
    int fail(const std::string &what)
    {
        err_ = "checkpoint: " + what + ": " + strerror(errno);
        return -1;
    }
    std::string final_, tmp_, err_;
    bool open_ = false;
};

// MANIFEST of format 1:  "pegasus_b200_checkpoint 1" | app_id | pidx | data_version | last_flushed_decree | last_seq | runs N |
// N lines "<level> <file> <bytes>" (oldest run first)
struct CheckpointManifest {
    long long app_id = 0, pidx = 0, data_version = -1, decree = -1, last_seq = 0;
    struct File { int level; std::string name; long long bytes; };
    std::vector<File> files;
    std::string str() const
    {
        std::string m = "pegasus_b200_checkpoint 1\n";
        m += "app_id " + std::to_string(app_id) + "\npidx " + std::to_string(pidx) + "\ndata_version " + std::to_string(data_version) + "\n";
        m += "last_flushed_decree " + std::to_string(decree) + "\nlast_seq " + std::to_string(last_seq) + "\nruns " + std::to_string(files.size()) + "\n";
        for (const File &f : files) m += std::to_string(f.level) + " " + f.name + " " + std::to_string(f.bytes) + "\n";
        return m;
    }
    // false = not a manifest of this format (damaged, truncated, absurd counts)
    bool parse(const std::string &text)
    {
        std::istringstream in(text);
        std::string word, k[6];
        int version = 0;
        long long nruns = -1;
        in >> word >> version;
        if (!in || word != "pegasus_b200_checkpoint" || version != 1) return false;
        in >> k[0] >> app_id >> k[1] >> pidx >> k[2] >> data_version >> k[3] >> decree >> k[4] >> last_seq >> k[5] >> nruns;
        if (!in || k[0] != "app_id" || k[1] != "pidx" || k[2] != "data_version" || k[3] != "last_flushed_decree" || k[4] != "last_seq" || k[5] != "runs")
            return false;
        if (nruns < 0 || nruns > 65536 || decree < 0 || last_seq < 0) return false;
        files.clear();
        for (long long i = 0; i < nruns; i++) {
            File f{};
            in >> f.level >> f.name >> f.bytes;
            if (!in || f.level < 0 || f.level > 64 || f.bytes < 0 || f.name.empty() || f.name.find('/') != std::string::npos || f.name == "." || f.name == "..")
                return false;
            files.push_back(f);
        }
        return true;
    }
};

} // namespace pgs
