// stage1_files.cuh -- Stage I straight from files (host side; included by kssd_b200.cu).
//
// Replaces the file loop of run_stageI (reference command_dist.c:277-312) together with the popen("zcat -fc") decode of
// fasta2co / fastq2co (iseq2comem.c:187-200, :283-290): reader threads pull files in input order and either read() plain
// files straight into a pinned staging buffer (their size is known, so their place in the batch is assigned before the
// read) or inflate .gz files with zlib into a private buffer that is copied to its place once its size is known.  A
// batch is closed when the next file would not fit; a GPU thread copies it to the device and sketches it while the
// readers fill the second staging buffer.  Results of all batches are appended in file order.
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>

struct kssd_stage1 {
    int n_files = 0, n_comp = 1;
    std::vector<std::vector<uint32_t>> ids;      // per component, files concatenated in input order
    std::vector<std::vector<uint16_t>> abund;    // per component (abundance mode)
    std::vector<std::vector<uint64_t>> index;    // per component, n_files + 1
    std::vector<int32_t> status;                 // per file: 0 / KSSD_E_*
    std::vector<uint64_t> file_bytes;            // decoded size of every file
    double read_s = 0, gpu_s = 0, total_s = 0;
    uint64_t bytes = 0;
    int batches = 0;
};

namespace stage1 {

using clk = std::chrono::steady_clock;
static inline double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

static bool has_gz_magic(const char *path)
{
    unsigned char m[2] = {0, 0};
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    const ssize_t n = read(fd, m, 2);
    close(fd);
    return n == 2 && m[0] == 0x1f && m[1] == 0x8b;
}

struct Task {                 // one file
    int file = -1;
    bool gz = false;
    uint64_t size = 0;        // decoded bytes (plain: from stat, gz: after inflate)
    uint8_t *priv = nullptr;  // gz: inflated data
    uint8_t *dest = nullptr;  // place in the staging buffer
    bool decoded = false, placed = false, failed = false;
};

struct Batch {
    int staging = 0;
    std::vector<int> files;
    std::vector<uint64_t> goff, glen;
    uint64_t bytes = 0;
};

// a minimal pool: jobs run in submission order over n threads
class Pool {
  public:
    explicit Pool(int n)
    {
        for (int i = 0; i < n; i++) th_.emplace_back([this] { run(); });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void submit(std::function<void()> f)
    {
        {
            std::lock_guard<std::mutex> l(m_);
            q_.push_back(std::move(f));
        }
        cv_.notify_one();
    }

  private:
    void run()
    {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    std::vector<std::thread> th_;
    bool stop_ = false;
};

static bool read_plain(const char *path, uint8_t *dst, uint64_t size)
{
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    uint64_t got = 0;
    while (got < size) {
        const ssize_t n = read(fd, dst + got, (size_t)std::min<uint64_t>(size - got, 1ull << 30));
        if (n <= 0) break;
        got += (uint64_t)n;
    }
    close(fd);
    return got == size;
}

// `-P <cmd>` of the reference (popen("<cmd> <file>"), iseq2comem.c:195-199): the command's stdout is the file content
static bool pipe_file(const char *cmd, const char *path, uint8_t **out, uint64_t *size)
{
    std::string line = std::string(cmd) + " " + path;
    FILE *fp = popen(line.c_str(), "r");
    if (!fp) return false;
    uint64_t cap = 1ull << 24, n = 0;
    uint8_t *buf = (uint8_t *)malloc(cap);
    if (!buf) { pclose(fp); return false; }
    for (;;) {
        if (cap - n < (1u << 20)) {
            cap *= 2;
            uint8_t *nb = (uint8_t *)realloc(buf, cap);
            if (!nb) { free(buf); pclose(fp); return false; }
            buf = nb;
        }
        const size_t r = fread(buf + n, 1, cap - n, fp);
        if (r == 0) break;
        n += r;
    }
    const int st = pclose(fp);
    if (st != 0) { free(buf); return false; }
    *out = buf;
    *size = n;
    return true;
}

static bool inflate_file(const char *path, uint8_t **out, uint64_t *size)
{
    gzFile g = gzopen(path, "rb");
    if (!g) return false;
    gzbuffer(g, 1u << 20);
    uint64_t cap = 1ull << 24, n = 0;
    uint8_t *buf = (uint8_t *)malloc(cap);
    if (!buf) { gzclose(g); return false; }
    for (;;) {
        if (cap - n < (1u << 20)) {
            cap *= 2;
            uint8_t *nb = (uint8_t *)realloc(buf, cap);
            if (!nb) { free(buf); gzclose(g); return false; }
            buf = nb;
        }
        const int r = gzread(g, buf + n, (unsigned)std::min<uint64_t>(cap - n, 1u << 30));
        if (r < 0) { free(buf); gzclose(g); return false; }
        if (r == 0) break;
        n += (uint64_t)r;
    }
    gzclose(g);
    *out = buf;
    *size = n;
    return true;
}

}  // namespace stage1

extern "C" int kssd_stage1_files_ex(kssd_ctx_t *c, const char *const *paths, int n_files, const kssd_sketch_opts_t *opts, int n_threads,
                                    size_t batch_bytes, const char *pipecmd, kssd_stage1_t **out);

extern "C" int kssd_stage1_files(kssd_ctx_t *c, const char *const *paths, int n_files, const kssd_sketch_opts_t *opts, int n_threads,
                                 size_t batch_bytes, kssd_stage1_t **out)
{
    return kssd_stage1_files_ex(c, paths, n_files, opts, n_threads, batch_bytes, nullptr, out);
}

extern "C" int kssd_stage1_files_ex(kssd_ctx_t *c, const char *const *paths, int n_files, const kssd_sketch_opts_t *opts, int n_threads,
                                    size_t batch_bytes, const char *pipecmd, kssd_stage1_t **out)
{
    using namespace stage1;
    if (!c || !paths || !out || n_files <= 0) return fail(KSSD_E_INVAL, "kssd_stage1_files: bad argument");
    const int mode = opts ? opts->mode : KSSD_MODE_FASTA;
    if (mode == KSSD_MODE_BYREAD) return fail(KSSD_E_INVAL, "kssd_stage1_files: --byread goes through kssd_sketch_batch_*");
    CU(cudaSetDevice(c->device));
    const int nt = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    if (batch_bytes == 0) batch_bytes = 1ull << 30;
    const auto t_start = clk::now();

    std::vector<Task> tasks(n_files);
    uint64_t biggest_plain = 0;
    for (int i = 0; i < n_files; i++) {
        tasks[i].file = i;
        struct stat st;
        if (!paths[i] || stat(paths[i], &st) != 0) return fail(KSSD_E_INVAL, "kssd_stage1_files: cannot stat %s", paths[i] ? paths[i] : "(null)");
        tasks[i].gz = (pipecmd && pipecmd[0]) || has_gz_magic(paths[i]);     // "gz" = size unknown until decoded
        if (!tasks[i].gz) { tasks[i].size = (uint64_t)st.st_size; biggest_plain = std::max(biggest_plain, tasks[i].size); }
    }
    // two pinned staging buffers, kept in the context between calls (pinning a GiB costs about as much as reading it);
    // a file larger than a batch gets a batch of its own (the buffers grow on demand)
    uint64_t *stag_cap = c->stag_cap;
    uint8_t **stag = c->stag;
    auto ensure_staging = [&](int b, uint64_t bytes) -> bool {
        if (bytes <= stag_cap[b]) return true;
        if (stag[b]) cudaFreeHost(stag[b]);
        stag[b] = nullptr;
        stag_cap[b] = 0;
        if (cudaHostAlloc((void **)&stag[b], bytes, cudaHostAllocDefault) != cudaSuccess) return false;
        stag_cap[b] = bytes;
        return true;
    };
    uint64_t expect = 0;                              // decoded bytes to expect (gz: a guess, the buffers grow if it is low)
    for (int i = 0; i < n_files; i++) {
        struct stat st;
        stat(paths[i], &st);
        expect += tasks[i].gz ? 5 * (uint64_t)st.st_size : tasks[i].size;
    }
    batch_bytes = (size_t)std::max<uint64_t>(std::min<uint64_t>(batch_bytes, expect + 16ull * n_files), 1u << 20);
    const uint64_t first_cap = std::max<uint64_t>(batch_bytes, biggest_plain) + 4096;
    if (!ensure_staging(0, first_cap) || !ensure_staging(1, first_cap))
        return fail(KSSD_E_NOMEM, "kssd_stage1_files: cannot pin %llu bytes of staging memory", (unsigned long long)first_cap);

    kssd_stage1 *R = new kssd_stage1();
    R->n_files = n_files;
    R->n_comp = c->info.component_num;
    R->ids.resize(R->n_comp);
    R->abund.resize(R->n_comp);
    R->index.assign(R->n_comp, std::vector<uint64_t>(1, 0));
    R->status.assign(n_files, 0);
    R->file_bytes.assign(n_files, 0);

    std::mutex m;
    std::condition_variable cv;
    uint64_t inflight_priv = 0;                       // bytes of inflated-but-unplaced data (back-pressure on the gz decoders)
    const uint64_t priv_limit = std::max<uint64_t>(4 * (uint64_t)batch_bytes, 1ull << 30);
    bool staging_free[2] = {true, true};
    std::deque<Batch> gpu_q;
    bool gpu_stop = false, aborting = false;
    int gpu_rc = KSSD_OK;
    std::string gpu_err;
    double read_busy = 0;

    // ---- GPU thread: one batch at a time, results appended in batch (= file) order
    std::thread gpu([&] {
        cudaSetDevice(c->device);
        for (;;) {
            Batch b;
            {
                std::unique_lock<std::mutex> l(m);
                cv.wait(l, [&] { return gpu_stop || !gpu_q.empty(); });
                if (gpu_q.empty()) return;
                b = std::move(gpu_q.front());
                gpu_q.pop_front();
            }
            const auto t0 = clk::now();
            int rc = gpu_rc;
            if (rc == KSSD_OK && !b.files.empty()) {
                kssd_sketch_t *sk = nullptr;
                rc = kssd_sketch_batch_host(c, stag[b.staging], b.bytes, b.goff.data(), b.glen.data(), (int)b.files.size(), opts, &sk);
                if (rc == KSSD_OK) {
                    const int nf = (int)b.files.size();
                    std::vector<uint64_t> ix(nf + 1);
                    for (int cc = 0; cc < R->n_comp && rc == KSSD_OK; cc++) {
                        const int64_t n = kssd_sketch_count(sk, cc);
                        const size_t base = R->ids[cc].size();
                        R->ids[cc].resize(base + (size_t)n);
                        const bool ab = mode == KSSD_MODE_FASTQ_ABUND;
                        if (ab) R->abund[cc].resize(base + (size_t)n);
                        rc = kssd_sketch_fetch(sk, cc, n ? R->ids[cc].data() + base : nullptr, ix.data(), ab && n ? R->abund[cc].data() + base : nullptr, nullptr);
                        const uint64_t off = R->index[cc].back();
                        for (int f = 0; f < nf; f++) R->index[cc].push_back(off + ix[f + 1]);
                    }
                    std::vector<int32_t> stt(nf);
                    if (rc == KSSD_OK) rc = kssd_sketch_status(sk, stt.data());
                    for (int f = 0; f < nf; f++) R->status[b.files[f]] = stt[f];
                    kssd_sketch_free(sk);
                }
                if (rc != KSSD_OK) gpu_err = kssd_last_error();
            }
            {
                std::lock_guard<std::mutex> l(m);
                if (rc != KSSD_OK && gpu_rc == KSSD_OK) gpu_rc = rc;
                R->gpu_s += secs(t0, clk::now());
                R->batches++;
                staging_free[b.staging] = true;
            }
            cv.notify_all();
        }
    });

    {
        // two pools: inflating (may wait for room) and placing (reads of plain files, copies of inflated ones) -- a
        // decoder that waits for room must never keep a placement from running
        Pool dec(nt), plc(nt);
        // gz files are inflated eagerly (bounded by priv_limit); plain files are read once their place is known
        for (int i = 0; i < n_files; i++)
            if (tasks[i].gz)
                dec.submit([&, i] {
                    {
                        std::unique_lock<std::mutex> l(m);
                        cv.wait(l, [&] { return aborting || inflight_priv < priv_limit; });
                        if (aborting) { tasks[i].failed = true; tasks[i].decoded = true; cv.notify_all(); return; }
                    }
                    const auto t0 = clk::now();
                    uint8_t *buf = nullptr;
                    uint64_t sz = 0;
                    const bool ok = (pipecmd && pipecmd[0]) ? pipe_file(pipecmd, paths[i], &buf, &sz) : inflate_file(paths[i], &buf, &sz);
                    std::lock_guard<std::mutex> l(m);
                    tasks[i].priv = buf; tasks[i].size = sz; tasks[i].failed = !ok; tasks[i].decoded = true;
                    inflight_priv += sz;
                    read_busy += secs(t0, clk::now());
                    cv.notify_all();
                });
        // ---- coordinator: files in input order -> batches
        int cur = -1;
        Batch B;
        int outstanding = 0;                              // placements of the open batch still running
        bool failed = false;
        auto open_batch = [&] {
            std::unique_lock<std::mutex> l(m);
            cv.wait(l, [&] { return staging_free[0] || staging_free[1]; });
            cur = staging_free[0] ? 0 : 1;
            staging_free[cur] = false;
            B = Batch();
            B.staging = cur;
        };
        // A batch goes to the GPU only if every file of it landed in the staging buffer: a file that stat() accepted but
        // open() / read() rejected (permissions, truncation after stat, EIO) must fail the call like the reference's
        // "eof or fread error" does -- never be sketched from whatever the buffer held.
        auto close_batch = [&] {
            bool bad = false;
            {
                std::unique_lock<std::mutex> l(m);
                cv.wait(l, [&] { return outstanding == 0; });
                for (int f : B.files) bad |= tasks[f].failed;
                if (!bad) gpu_q.push_back(std::move(B));
                else staging_free[cur] = true;
            }
            if (bad) failed = true;
            cv.notify_all();
            cur = -1;
        };
        for (int i = 0; i < n_files && !failed; i++) {
            Task &T = tasks[i];
            if (T.gz) {
                std::unique_lock<std::mutex> l(m);
                cv.wait(l, [&] { return T.decoded; });
            }
            if (T.failed) { failed = true; break; }
            const uint64_t need = (T.size + 15) & ~15ull;
            if (cur >= 0 && B.bytes + need > batch_bytes && !B.files.empty()) close_batch();
            if (cur < 0) open_batch();
            if (B.bytes + need + 4096 > stag_cap[cur]) {       // one oversized file: this batch is empty, grow its buffer
                if (!ensure_staging(cur, B.bytes + need + 4096)) { failed = true; break; }
            }
            T.dest = stag[cur] + B.bytes;
            B.files.push_back(i);
            B.goff.push_back(B.bytes);
            B.glen.push_back(T.size);
            B.bytes += need;
            R->file_bytes[i] = T.size;
            R->bytes += T.size;
            {
                std::lock_guard<std::mutex> l(m);
                outstanding++;
            }
            plc.submit([&, i] {
                Task &t = tasks[i];
                const auto t0 = clk::now();
                bool ok = true;
                if (t.gz) { memcpy(t.dest, t.priv, t.size); free(t.priv); t.priv = nullptr; }
                else ok = read_plain(paths[i], t.dest, t.size);
                memset(t.dest + t.size, '\n', ((t.size + 15) & ~15ull) - t.size);
                std::lock_guard<std::mutex> l(m);
                if (t.gz) inflight_priv -= t.size;
                t.placed = true; t.failed = !ok;
                outstanding--;
                read_busy += secs(t0, clk::now());
                cv.notify_all();
            });
        }
        if (cur >= 0) {
            if (B.files.empty() || failed) {
                std::unique_lock<std::mutex> l(m);
                cv.wait(l, [&] { return outstanding == 0; });
                staging_free[cur] = true;
            } else close_batch();
        }
        {   // drain
            std::unique_lock<std::mutex> l(m);
            cv.wait(l, [&] { return gpu_q.empty() && staging_free[0] && staging_free[1]; });
            gpu_stop = true;
            aborting = true;                              // decoders still queued after a failure give up at the gate
        }
        cv.notify_all();
        gpu.join();
        // the pool destructors join the readers
        if (failed) gpu_rc = gpu_rc == KSSD_OK ? KSSD_E_INVAL : gpu_rc;
    }
    for (auto &t : tasks) {
        if (t.priv) free(t.priv);
        if (t.failed) {
            if (gpu_err.empty()) gpu_err = std::string("cannot read ") + paths[t.file];
            if (gpu_rc == KSSD_OK) gpu_rc = KSSD_E_INVAL;     // no read failure is ever silent
        }
    }
    R->read_s = read_busy / nt;
    R->total_s = secs(t_start, clk::now());
    if (gpu_rc != KSSD_OK) {
        delete R;
        return fail(gpu_rc, "kssd_stage1_files: %s", gpu_err.c_str());
    }
    *out = R;
    return KSSD_OK;
}

extern "C" int64_t kssd_stage1_count(const kssd_stage1_t *s, int comp)
{
    if (!s || comp < 0 || comp >= s->n_comp) return fail(KSSD_E_INVAL, "kssd_stage1_count: bad component");
    return (int64_t)s->ids[comp].size();
}

extern "C" int kssd_stage1_fetch(const kssd_stage1_t *s, int comp, uint32_t *ids, uint64_t *index, uint16_t *abund)
{
    if (!s || comp < 0 || comp >= s->n_comp) return fail(KSSD_E_INVAL, "kssd_stage1_fetch: bad component");
    if (ids && !s->ids[comp].empty()) memcpy(ids, s->ids[comp].data(), s->ids[comp].size() * 4);
    if (index) memcpy(index, s->index[comp].data(), s->index[comp].size() * 8);
    if (abund && !s->abund[comp].empty()) memcpy(abund, s->abund[comp].data(), s->abund[comp].size() * 2);
    return KSSD_OK;
}

extern "C" int kssd_stage1_status(const kssd_stage1_t *s, int32_t *status_out)
{
    if (!s || !status_out) return fail(KSSD_E_INVAL, "kssd_stage1_status: null");
    memcpy(status_out, s->status.data(), s->status.size() * 4);
    return KSSD_OK;
}

extern "C" int kssd_stage1_timing(const kssd_stage1_t *s, double *read_s, double *gpu_s, double *total_s, uint64_t *bytes, int *batches)
{
    if (!s) return fail(KSSD_E_INVAL, "kssd_stage1_timing: null");
    if (read_s) *read_s = s->read_s;
    if (gpu_s) *gpu_s = s->gpu_s;
    if (total_s) *total_s = s->total_s;
    if (bytes) *bytes = s->bytes;
    if (batches) *batches = s->batches;
    return KSSD_OK;
}

extern "C" void kssd_stage1_free(kssd_stage1_t *s) { delete s; }
