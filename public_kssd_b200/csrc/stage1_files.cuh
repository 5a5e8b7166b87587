// stage1_files.cuh -- Stage I straight from files (host side; included by kssd_b200.cu).
//
// Replaces the file loop of run_stageI (reference command_dist.c:277-312) together with the popen("zcat -fc") decode of
// fasta2co / fastq2co (iseq2comem.c:187-200, :283-290): reader threads pull files in input order and either read() plain
// files straight into a pinned staging buffer (their size is known, so their place in the batch is assigned before the
// read) or inflate .gz files with zlib into a private buffer that is copied to its place once its size is known.  A
// batch is closed when the next file would not fit; a GPU thread copies it to the device and sketches it while the
// readers fill the second staging buffer.  Results of all batches are appended in file order.
// A call with 640 or more .gz files takes the other road (run_gz_gpu below): the files go to the device as they are and are
// inflated there, one file per warp (inflate.cuh).
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "inflate.cuh"

struct kssd_stage1 {
    int n_files = 0, n_comp = 1;
    std::vector<std::vector<uint32_t>> ids;      // per component, files concatenated in input order
    std::vector<std::vector<uint16_t>> abund;    // per component (abundance mode)
    std::vector<std::vector<uint64_t>> index;    // per component, n_files + 1
    std::vector<int32_t> status;                 // per file: 0 / KSSD_E_*
    std::vector<uint64_t> file_bytes;            // decoded size of every file
    double read_s = 0, gpu_s = 0, total_s = 0, gz_gpu_s = 0;      // gz_gpu_s: H2D of the compressed bytes + inflate on the GPU (part of gpu_s)
    bool gz_on_gpu = false;
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

// the sketch of one batch appended to the call's result, files in batch order
static int append_results(kssd_stage1 *R, kssd_sketch_t *sk, const std::vector<int> &files, int mode)
{
    const int nf = (int)files.size();
    std::vector<uint64_t> ix(nf + 1);
    int rc = KSSD_OK;
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
    for (int f = 0; f < nf; f++) R->status[files[f]] = stt[f];
    return rc;
}

// ---- .gz decoded on the GPU (inflate.cuh) -------------------------------------------------------------------------------
// The files of a batch are read as they are (compressed) into the pinned staging buffer and copied to the device; one thread per
// file inflates into the batch's text buffer (every file's decoded size is its gzip ISIZE), plain files of the batch are copied
// device to device, and the batch is sketched where it lies.  A batch is sized by DECODED bytes (16 GiB unless
// KSSD_GZ_BATCH_BYTES says otherwise, the batches of a call equal in size): the decoder's parallelism is the number of files in flight.  Any file the kernel cannot
// finish in its ISIZE bytes (several gzip members, corrupt data, CRC mismatch) sets *fall_back: the caller redoes the call with
// zlib on the host, which also produces the error message if the file is really broken.
struct GzFile { bool gz = false; uint64_t csize = 0, dsize = 0; };

static bool gz_isize(const char *path, uint64_t fsize, uint64_t *isize)
{
    if (fsize < 18) return false;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    unsigned char t[4];
    const bool ok = pread(fd, t, 4, (off_t)(fsize - 4)) == 4;
    close(fd);
    *isize = (uint64_t)t[0] | ((uint64_t)t[1] << 8) | ((uint64_t)t[2] << 16) | ((uint64_t)t[3] << 24);
    return ok;
}

static int run_gz_gpu(kssd_ctx_t *c, const char *const *paths, int n_files, const kssd_sketch_opts_t *opts, int nt, const std::vector<GzFile> &F,
                      kssd_stage1 *R, bool *fall_back)
{
    const int mode = opts ? opts->mode : KSSD_MODE_FASTA;
    const char *eb = getenv("KSSD_GZ_BATCH_BYTES");
    const uint64_t cap_dec = eb ? std::max<uint64_t>(strtoull(eb, nullptr, 10), 1u << 20) : (16ull << 30);
    const bool check_crc = !getenv("KSSD_GZ_NOCRC");
    std::vector<std::pair<int, int>> batches;
    {
        // batches of equal decoded size (a short last batch would leave most of the GPU's decoder threads idle for a whole pass)
        uint64_t total = 0;
        for (int i = 0; i < n_files; i++) total += (F[i].dsize + 15) & ~15ull;
        const uint64_t nb = std::max<uint64_t>(1, (total + cap_dec - 1) / cap_dec), per = (total + nb - 1) / nb;
        uint64_t acc = 0, left = total;                   // left: bytes of the files not placed yet
        int lo = 0;
        for (int i = 0; i < n_files; i++) {
            const uint64_t need = (F[i].dsize + 15) & ~15ull;
            if (i > lo && (acc + need > cap_dec || (acc >= per && left > 0))) { batches.push_back({lo, i}); lo = i; acc = 0; }
            acc += need;
            left -= need;
        }
        batches.push_back({lo, n_files});
    }
    auto staged = [&](int i) { return (F[i].csize + 31) & ~15ull; };      // >= 16 zero bytes behind every file
    auto ensure_staging = [&](int b, uint64_t bytes) -> bool {
        if (bytes <= c->stag_cap[b]) return true;
        if (c->stag[b]) cudaFreeHost(c->stag[b]);
        c->stag[b] = nullptr;
        c->stag_cap[b] = 0;
        if (cudaHostAlloc((void **)&c->stag[b], bytes, cudaHostAllocDefault) != cudaSuccess) return false;
        c->stag_cap[b] = bytes;
        return true;
    };
    double read_busy = 0;
    std::mutex m;
    // reads batch k into staging buffer k & 1 on nt threads; joined by the returned thread
    std::vector<char> read_ok(batches.size(), 1);
    auto start_read = [&](size_t k) -> std::thread {
        return std::thread([&, k] {
            cudaSetDevice(c->device);                     // (this thread may pin the staging buffer: on the context's device, not on device 0)
            const int lo = batches[k].first, hi = batches[k].second;
            uint64_t bytes = 0;
            std::vector<uint64_t> soff(hi - lo);
            for (int i = lo; i < hi; i++) { soff[i - lo] = bytes; bytes += staged(i); }
            if (!ensure_staging((int)(k & 1), bytes + 4096)) { read_ok[k] = 0; return; }
            uint8_t *base = c->stag[k & 1];
            std::atomic<int> next{lo};
            std::atomic<bool> ok{true};
            std::vector<std::thread> th;
            const int workers = std::max(1, std::min(nt, hi - lo));
            for (int w = 0; w < workers; w++)
                th.emplace_back([&] {
                    const auto t0 = clk::now();
                    for (;;) {
                        const int i = next.fetch_add(1);
                        if (i >= hi) break;
                        uint8_t *dst = base + soff[i - lo];
                        if (!read_plain(paths[i], dst, F[i].csize)) ok = false;
                        memset(dst + F[i].csize, 0, staged(i) - F[i].csize);
                    }
                    std::lock_guard<std::mutex> l(m);
                    read_busy += secs(t0, clk::now());
                });
            for (auto &t : th) t.join();
            if (!ok) read_ok[k] = 0;
        });
    };
    std::thread rd = start_read(0);
    for (size_t k = 0; k < batches.size(); k++) {
        if (rd.joinable()) rd.join();
        if (!read_ok[k]) return fail(KSSD_E_INVAL, "kssd_stage1_files: cannot read a file of batch %zu", k);
        if (k + 1 < batches.size()) rd = start_read(k + 1);
        struct Joiner { std::thread *t; bool armed; ~Joiner() { if (armed && t->joinable()) t->join(); } } joiner{&rd, k + 1 < batches.size()};
        const auto t0 = clk::now();
        const int lo = batches[k].first, hi = batches[k].second, n = hi - lo;
        std::vector<uint64_t> soff(n), goff(n), glen(n);
        uint64_t sbytes = 0, tbytes = 0;
        for (int i = lo; i < hi; i++) {
            soff[i - lo] = sbytes; sbytes += staged(i);
            goff[i - lo] = tbytes; glen[i - lo] = F[i].dsize; tbytes += (F[i].dsize + 15) & ~15ull;
        }
        if (c->gzin.ensure(sbytes + 64) != cudaSuccess || c->gztext.ensure(tbytes + 1024) != cudaSuccess) {
            cudaGetLastError();                           // no room for the batch on the device: the host path needs far less
            c->gzin.release();
            c->gztext.release();
            *fall_back = true;
            return KSSD_OK;
        }
        uint8_t *d_in = c->gzin.as<uint8_t>(), *d_text = c->gztext.as<uint8_t>();
        CU(cudaMemcpyAsync(d_in, c->stag[k & 1], sbytes, cudaMemcpyHostToDevice, c->stream));
        std::vector<int> order;                           // gz files, largest first: the ticket order
        for (int i = lo; i < hi; i++) {
            if (F[i].gz) order.push_back(i);
            else if (F[i].dsize) CU(cudaMemcpyAsync(d_text + goff[i - lo], d_in + soff[i - lo], F[i].dsize, cudaMemcpyDeviceToDevice, c->stream));
        }
        std::sort(order.begin(), order.end(), [&](int a, int b) { return F[a].csize != F[b].csize ? F[a].csize > F[b].csize : a < b; });
        const uint32_t nj = (uint32_t)order.size();
        std::vector<kssd::gz::Job> jobs(nj);
        for (uint32_t j = 0; j < nj; j++) {
            const int i = order[j];
            jobs[j].in_off = soff[i - lo]; jobs[j].in_len = F[i].csize; jobs[j].out_off = goff[i - lo]; jobs[j].out_cap = F[i].dsize;
        }
        StreamScratch scr(c->stream);
        std::vector<kssd::gz::Result> res(nj);
        uint64_t *d_goff = static_cast<uint64_t *>(scr.alloc((size_t)n * 8)), *d_glen = static_cast<uint64_t *>(scr.alloc((size_t)n * 8));
        if (!d_goff || !d_glen) return fail(KSSD_E_NOMEM, "kssd_stage1_files: out of device memory");
        CU(cudaMemcpyAsync(d_goff, goff.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_glen, glen.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        if (nj) {
            kssd::gz::Job *d_jobs = static_cast<kssd::gz::Job *>(scr.alloc((size_t)nj * sizeof(kssd::gz::Job)));
            kssd::gz::Result *d_res = static_cast<kssd::gz::Result *>(scr.alloc((size_t)nj * sizeof(kssd::gz::Result)));
            uint32_t *d_ticket = static_cast<uint32_t *>(scr.alloc(4));
            if (!d_jobs || !d_res || !d_ticket) return fail(KSSD_E_NOMEM, "kssd_stage1_files: out of device memory");
            CU(cudaMemcpyAsync(d_jobs, jobs.data(), (size_t)nj * sizeof(kssd::gz::Job), cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemsetAsync(d_ticket, 0, 4, c->stream));
            // one file per warp (two files in a warp would only take turns); as many warps as the SMs hold (32 each), the rest of
            // the files pulled by ticket as warps finish
            const uint32_t per_sm = std::min<uint32_t>(32u, (uint32_t)((227u * 1024u) / (sizeof(kssd::gz::Tables) + 1024u)));
            const uint32_t active = 1;
            const uint32_t grid = std::min<uint32_t>((uint32_t)c->sm_count * per_sm, nj);
            const size_t smem = (size_t)active * sizeof(kssd::gz::Tables);
            CU(cudaFuncSetAttribute(kssd::gz::gunzip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kssd::gz::gunzip_kernel<<<grid, 32, smem, c->stream>>>(d_in, d_text, d_jobs, d_res, nj, active, d_ticket, check_crc ? 1 : 0);
            LAUNCHED(1);
            CU(cudaMemcpyAsync(res.data(), d_res, (size_t)nj * sizeof(kssd::gz::Result), cudaMemcpyDeviceToHost, c->stream));
        }
        kssd::gz::pad_text_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, c->stream>>>(d_text, d_goff, d_glen, (uint32_t)n);
        LAUNCHED(1);
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        for (uint32_t j = 0; j < nj; j++)
            if (res[j].status != kssd::gz::kOk || res[j].out_len != F[order[j]].dsize) { *fall_back = true; return KSSD_OK; }
        R->gz_gpu_s += secs(t0, clk::now());
        kssd_sketch_t *sk = nullptr;
        int rc = kssd_sketch_batch_dev(c, d_text, tbytes, goff.data(), glen.data(), n, opts, &sk);
        if (rc != KSSD_OK) return rc;
        std::vector<int> files(n);
        for (int i = 0; i < n; i++) files[i] = lo + i;
        rc = append_results(R, sk, files, mode);
        kssd_sketch_free(sk);
        if (rc != KSSD_OK) return rc;
        for (int i = lo; i < hi; i++) { R->file_bytes[i] = F[i].dsize; R->bytes += F[i].dsize; }
        R->gpu_s += secs(t0, clk::now());
        R->batches++;
    }
    R->read_s = read_busy / nt;
    return KSSD_OK;
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

    // enough .gz files to keep the GPU's decoder busy (or KSSD_GZ_GPU=1): compressed bytes cross PCIe, inflate runs on the device.
    // One stream takes the GPU thread ~0.13 s per MB of text however many run beside it (up to ~1,800 at once), the host's cores
    // inflate ~4 GB/s together: measured break-even at about 600 files of 5 MB (profiles/r2_gz_summary.md)
    {
        int n_gz = 0;
        for (int i = 0; i < n_files; i++) n_gz += tasks[i].gz ? 1 : 0;
        const char *eg = getenv("KSSD_GZ_GPU");
        bool use = !(pipecmd && pipecmd[0]) && n_gz > 0 && (eg ? atoi(eg) != 0 : n_gz >= 640);
        std::vector<GzFile> F(n_files);
        for (int i = 0; i < n_files && use; i++) {
            struct stat st;
            stat(paths[i], &st);
            F[i].gz = tasks[i].gz;
            F[i].csize = (uint64_t)st.st_size;
            F[i].dsize = F[i].csize;
            // (a compressed file much larger than its ISIZE: more than 4 GiB of text, or several members -- zlib's business)
            if (F[i].gz && (!gz_isize(paths[i], F[i].csize, &F[i].dsize) || F[i].dsize > (1ull << 30) || F[i].csize > F[i].dsize + F[i].dsize / 1000 + 1024)) use = false;
        }
        if (use) {
            bool fall_back = false;
            const int rc = run_gz_gpu(c, paths, n_files, opts, nt, F, R, &fall_back);
            if (c->gztext.cap > (1ull << 30)) { c->gzin.release(); c->gztext.release(); }      // (a big batch's buffers are not kept: Stage II / III want the memory)
            if (rc != KSSD_OK) { delete R; return rc; }
            if (!fall_back) {
                R->gz_on_gpu = true;
                R->total_s = secs(t_start, clk::now());
                *out = R;
                return KSSD_OK;
            }
            // some file did not decode into its ISIZE bytes: start over with zlib on the host
            for (int cc = 0; cc < R->n_comp; cc++) { R->ids[cc].clear(); R->abund[cc].clear(); R->index[cc].assign(1, 0); }
            R->status.assign(n_files, 0);
            R->file_bytes.assign(n_files, 0);
            R->bytes = 0; R->batches = 0; R->gpu_s = 0; R->read_s = 0; R->gz_gpu_s = 0;
        }
    }

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
                    rc = append_results(R, sk, b.files, mode);
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

extern "C" int kssd_stage1_gz_info(const kssd_stage1_t *s, int *on_gpu, double *gz_gpu_s)
{
    if (!s) return fail(KSSD_E_INVAL, "kssd_stage1_gz_info: null");
    if (on_gpu) *on_gpu = s->gz_on_gpu ? 1 : 0;
    if (gz_gpu_s) *gz_gpu_s = s->gz_gpu_s;
    return KSSD_OK;
}

// the GPU's gzip decoder (inflate.cuh) run on the host, for the CPU test suite: 0 or one of its negative codes
extern "C" int kssd_gunzip_host(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len)
{
    if (!in || !out_len || (cap && !out)) return fail(KSSD_E_INVAL, "kssd_gunzip_host: null");
    if (cap > 0xffffffffull) return fail(KSSD_E_INVAL, "kssd_gunzip_host: output larger than 4 GiB");
    std::vector<uint8_t> padded(n + kssd::gz::kInPad, 0);             // the decoder reads whole words: zeros behind the data
    memcpy(padded.data(), in, n);
    std::vector<uint64_t> aligned((cap + 7) / 8 + 1);                  // and checks the CRC from an 8-byte aligned text
    std::unique_ptr<kssd::gz::Tables> T(new kssd::gz::Tables);
    uint32_t len = 0;
    const int rc = kssd::gz::gunzip(padded.data(), n, reinterpret_cast<uint8_t *>(aligned.data()), (uint32_t)cap, *T, kssd::gz::h_crc8, &len);
    if (rc == 0 && len) memcpy(out, aligned.data(), len);
    *out_len = (size_t)len;
    return rc;
}

extern "C" void kssd_stage1_free(kssd_stage1_t *s) { delete s; }
