// See refio.h.  Behaviour restated from the reference's src/misc.c and src/freadmat.c
// (citations per function); written from the observed file/stdout conventions, not
// copied: std::vector storage, no global scratch, C++ I/O helpers.
#include "refio.h"

#include <algorithm>
#include <atomic>
#include <sys/stat.h>
#include <unistd.h>
#include <memory>
#include <thread>
#include <charconv>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <ctime>

namespace refio {

int debugmode = 0;

void die(const char *fmt, ...) {
    // several fits of the thread pool can fail in the same round: the first caller reports and ends the
    // process, the others wait for that (the reference, single-threaded, prints exactly one line)
    static std::atomic<bool> dying{false};
    if (dying.exchange(true))
        for (;;) pause();
    char msg[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof msg, fmt, ap);
    va_end(ap);
    printf("\n#! Terminated: %s.\n", msg);
    // The reference is single-threaded and calls exit(1) (src/misc.c:40-49).  Here the message may come from a
    // worker thread of the fit pool: exit() would run the static destructors -- the pool joining its own thread --
    // and abort.  Flush every open stream (the stdout protocol line, partially written output files: what exit()
    // would have done for them) and leave without unwinding; the status is the reference's.
    fflush(nullptr);
    _exit(1);
}

void warn(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    printf("\n#! Warning: ");
    vprintf(fmt, ap);
    printf(". ");
    va_end(ap);
}

void dprint(int level, const char *fmt, ...) {
    if (debugmode < level) return;
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
}

double cpu_seconds() { return (double)clock() / CLOCKS_PER_SEC; }

// position of the '.' that starts an extension made of alphanumerics, or npos
static size_t ext_pos(const std::string &name) {
    int i = (int)name.size() - 1;
    while (i >= 0 && isalnum((unsigned char)name[i])) i--;
    if (i > 0 && name[i] == '.' && name[i - 1] != '/' && name[i - 1] != '.') return (size_t)i;
    return std::string::npos;
}
bool has_extension(const std::string &name) { return ext_pos(name) != std::string::npos; }
std::string strip_extension(const std::string &name) {
    size_t p = ext_pos(name);
    return p == std::string::npos ? name : name.substr(0, p);
}
FILE *open_file(const std::string &name, const char *ext, const char *mode, std::string *used) {
    std::string fname = has_extension(name) ? name : name + ext;
    FILE *fp = fopen(fname.c_str(), mode);
    if (!fp) die("cant open %s", fname.c_str());
    if (used) *used = fname;
    return fp;
}

// ------------------------------------------------------------------ ascii in
void skip_junk(FILE *fp) {
    for (;;) {
        int c;
        do c = getc(fp);
        while (isspace(c));
        if (c == '#') {
            while ((c = getc(fp)) != '\n' && c != EOF) {}
            if (c == EOF) return;
        } else {
            ungetc(c, fp);
            return;
        }
    }
}
void skip_line(FILE *fp) {
    for (;;) {
        int c = getc(fp);
        if (c == EOF || c == '\n') return;
        if (c == '\r') {
            c = getc(fp);
            if (c != '\n') ungetc(c, fp);
            return;
        }
    }
}
void skip_word(FILE *fp) {
    for (;;) {
        int c = getc(fp);
        if (c == EOF || isspace(c)) return;
    }
}
int read_int(FILE *fp) {
    int x;
    skip_junk(fp);
    if (fscanf(fp, "%d", &x) != 1) die("cant read int");
    return x;
}
double read_double(FILE *fp) {
    double x;
    skip_junk(fp);
    if (fscanf(fp, "%lf", &x) != 1) die("cant read double");
    return x;
}
std::string read_word(FILE *fp) {
    std::string w;
    skip_junk(fp);
    for (;;) {
        int c = getc(fp);
        if (c == EOF) break;
        if (isspace(c)) {
            ungetc(c, fp);
            break;
        }
        w.push_back((char)c);
        if (w.size() >= BUFSIZ) die("too long word");
    }
    return w;
}
std::vector<double> read_vec(FILE *fp, int expect) {
    int m = read_int(fp);
    if (expect > 0 && expect != m) die("size mismatch in vec");
    std::vector<double> a(m > 0 ? m : 0);
    for (int i = 0; i < m; i++) a[i] = read_double(fp);
    return a;
}
std::vector<int> read_ivec(FILE *fp, int expect) {
    int m = read_int(fp);
    if (expect > 0 && expect != m) die("size mismatch in ivec");
    std::vector<int> a(m > 0 ? m : 0);
    for (int i = 0; i < m; i++) a[i] = read_int(fp);
    return a;
}
std::vector<std::string> read_svec(FILE *fp) {
    int m = read_int(fp);
    std::vector<std::string> a;
    for (int i = 0; i < m; i++) a.push_back(read_word(fp));
    return a;
}
// The rest of a regular file in memory, with a '\n' sentinel behind it.  Every host thread reads its own
// slice with pread into an uninitialised buffer, so the first touch of the pages (the expensive part of a
// gigabyte-sized allocation) and the copy out of the page cache run on all cores.  The stream position of
// `fp` does not move.
namespace {
// isspace of the "C" locale as an inline test (the parsers below call it once per byte)
inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

int io_threads() {
    unsigned hw = std::thread::hardware_concurrency();
    // one process per GPU shares the host: torchrun exports LOCAL_WORLD_SIZE
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) hw /= (unsigned)std::max(1, atoi(e));
    return (int)std::max(1u, std::min(64u, hw));
}

struct FileTail {
    std::unique_ptr<char[]> buf;
    size_t bytes = 0;
    long pos = 0;
    bool load(FILE *fp) {
        struct stat st;
        const int fd = fileno(fp);
        if (fd < 0 || fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return false;
        pos = ftell(fp);
        if (pos < 0 || st.st_size <= pos) return false;
        bytes = (size_t)(st.st_size - pos);
        buf.reset(new (std::nothrow) char[bytes + 1]);
        if (!buf) return false;
        const int nthr = bytes < ((size_t)8 << 20) ? 1 : io_threads();
        std::vector<char> good(nthr, 1);
        auto work = [&](int t) {
            size_t lo = bytes / nthr * t, hi = t == nthr - 1 ? bytes : bytes / nthr * (t + 1);
            while (lo < hi) {
                const ssize_t r = pread(fd, buf.get() + lo, hi - lo, (off_t)pos + (off_t)lo);
                if (r <= 0) { good[t] = 0; return; }
                lo += (size_t)r;
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthr; t++) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
        for (int t = 0; t < nthr; t++)
            if (!good[t]) return false;
        buf[bytes] = '\n';
        return true;
    }
};
}  // namespace

// Parallel text parse of `count` doubles from a regular file (a K=1000 x n=10^5
// matrix is 2 GB of text, minutes of fscanf).  The rest of the file is read into memory, cut
// at line ends (a '#' comment never spans a line end), tokenised and converted with
// std::from_chars on all host threads, and copied into place on all host threads.  Anything the
// strict tokenizer does not understand (hex floats, trailing garbage, too few values) makes it
// give up WITHOUT consuming input, and the caller falls back to the fscanf loop, so error
// behaviour stays the reference's.
static bool read_doubles_parallel(FILE *fp, double *out, size_t count) {
    if (count < (1u << 16)) return false;
    FileTail file;
    if (!file.load(fp)) return false;
    const size_t bytes = file.bytes;
    const long pos = file.pos;
    const char *base = file.buf.get();
    const int nthr = io_threads();
    std::vector<size_t> cut(nthr + 1, bytes);
    cut[0] = 0;
    for (int t = 1; t < nthr; t++) {
        size_t c = std::max(cut[t - 1], bytes / nthr * t);
        const void *nl = c < bytes ? memchr(base + c, '\n', bytes - c) : nullptr;
        cut[t] = nl ? (size_t)(static_cast<const char *>(nl) - base) + 1 : bytes;
    }
    // tokens of [lo, hi): every value, or stop as soon as `limit` values were taken
    struct Part {
        std::vector<double> vals;
        std::vector<size_t> ends;  // only filled when asked for: offset just behind every token
        size_t end = 0;            // offset just behind the last token taken
        bool ok = true;
    };
    auto parse = [&](size_t lo, size_t hi, size_t limit, Part *p) {
        const char *c = base + lo, *e = base + hi;
        while (c < e && p->vals.size() < limit) {
            if (is_space((unsigned char)*c)) { c++; continue; }
            if (*c == '#') {
                const void *nl = memchr(c, '\n', (size_t)(e - c));
                c = nl ? static_cast<const char *>(nl) : e;
                continue;
            }
            const char *tok = c;
            while (c < e && !is_space((unsigned char)*c)) c++;
            const char *b = tok;
            if (*b == '+') b++;
            double v;
            auto r = std::from_chars(b, c, v);
            if (r.ec != std::errc() || r.ptr != c || (b + 1 < c && b[0] == '0' && (b[1] == 'x' || b[1] == 'X'))) {
                p->ok = false;
                return;
            }
            p->vals.push_back(v);
            p->end = (size_t)(c - base);
        }
    };
    std::vector<Part> parts(nthr);
    {
        std::vector<std::thread> th;
        auto work = [&](int t) {
            try {  // a value takes at least two bytes; typical files spend 15-25 on one
                parts[t].vals.reserve((cut[t + 1] - cut[t]) / 12 + 16);
            } catch (...) {
            }
            parse(cut[t], cut[t + 1], ~(size_t)0, &parts[t]);
        };
        for (int t = 1; t < nthr; t++) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
    }
    // the matrix may be followed by other data: take exactly `count` values in file order
    size_t have = 0, end = 0;
    bool ok = true;
    std::vector<size_t> first(nthr, 0), take(nthr, 0);
    for (int t = 0; t < nthr && have < count; t++) {
        if (!parts[t].ok && parts[t].vals.size() < count - have) { ok = false; break; }
        take[t] = std::min(parts[t].vals.size(), count - have);
        first[t] = have;
        if (take[t] == parts[t].vals.size()) {
            end = parts[t].end ? parts[t].end : end;
        } else {  // the matrix ends inside this part: find where, sequentially
            Part q;
            parse(cut[t], cut[t + 1], take[t], &q);
            if (!q.ok || q.vals.size() != take[t]) { ok = false; break; }
            end = q.end;
        }
        have += take[t];
    }
    if (!ok || have != count) {
        fseek(fp, pos, SEEK_SET);
        return false;
    }
    {
        std::vector<std::thread> th;
        auto copy = [&](int t) {
            if (take[t]) memcpy(out + first[t], parts[t].vals.data(), take[t] * sizeof(double));
        };
        for (int t = 1; t < nthr; t++) th.emplace_back(copy, t);
        copy(0);
        for (auto &x : th) x.join();
    }
    fseek(fp, pos + (long)end, SEEK_SET);
    return true;
}

std::vector<double> read_mat(FILE *fp, int *mp, int *np) {
    int m = read_int(fp), n = read_int(fp);
    if (*mp > 0 && *mp != m) die("size of rows mismatch in mat");
    if (*np > 0 && *np != n) die("size of columns mismatch in mat");
    std::vector<double> a;
    if ((long long)m * n == 0) return a;
    a.resize((size_t)m * n);
    if (!read_doubles_parallel(fp, a.data(), a.size()))
        for (size_t e = 0; e < a.size(); e++) a[e] = read_double(fp);
    *mp = m;
    *np = n;
    return a;
}

// ------------------------------------------------------------------ ascii out
static const int kPerLine = 5;
void write_vec(FILE *fp, const double *a, int m) {
    fprintf(fp, "#!VEC:\n%d\n", m);
    for (int i = 0, k = 0; i < m; i++) {
        if (++k > kPerLine) {
            putc('\n', fp);
            k = 1;
        }
        fprintf(fp, "%14.9g ", a[i]);
    }
    putc('\n', fp);
}
void write_ivec(FILE *fp, const int *a, int m) {
    fprintf(fp, "#!VEC:\n%d\n", m);
    for (int i = 0, k = 0; i < m; i++) {
        if (++k > kPerLine) {
            putc('\n', fp);
            k = 1;
        }
        fprintf(fp, "%10d ", a[i]);
    }
    putc('\n', fp);
}
void write_mat(FILE *fp, const double *a, int m, int n) {
    fprintf(fp, "#!MAT:\n%d %d\n", m, n);
    if (!a) return;
    for (int i = 0; i < m; i++) {
        fprintf(fp, "\n# row: %d\n", i);
        for (int j = 0, k = 0; j < n; j++) {
            if (++k > kPerLine) {
                putc('\n', fp);
                k = 1;
            }
            fprintf(fp, "%14.9g ", a[(size_t)i * n + j]);
        }
        putc('\n', fp);
    }
}

// ------------------------------------------------------------------ binary
int read_bint(FILE *fp) {
    int x;
    if (fread(&x, sizeof(int), 1, fp) != 1) die("cant read binary int");
    return x;
}
std::vector<double> read_bvec(FILE *fp, int expect) {
    int m = read_bint(fp);
    if (expect > 0 && expect != m) die("size mismatch in bvec");
    std::vector<double> a(m > 0 ? m : 0);
    if (m > 0 && fread(a.data(), sizeof(double), m, fp) != (size_t)m) die("cant read bvec");
    return a;
}
std::vector<int> read_bivec(FILE *fp, int expect) {
    int m = read_bint(fp);
    if (expect > 0 && expect != m) die("size mismatch in bivec");
    std::vector<int> a(m > 0 ? m : 0);
    if (m > 0 && fread(a.data(), sizeof(int), m, fp) != (size_t)m) die("cant read bivec");
    return a;
}
std::vector<double> read_bmat(FILE *fp, int *mp, int *np) {
    int m = read_bint(fp), n = read_bint(fp);
    if (*mp > 0 && *mp != m) die("size of rows mismatch in bmat");
    if (*np > 0 && *np != n) die("size of columns mismatch in bmat");
    std::vector<double> a((size_t)m * n);
    if (!a.empty() && fread(a.data(), sizeof(double), a.size(), fp) != a.size()) die("cant read bmat");
    *mp = m;
    *np = n;
    return a;
}
void write_bint(FILE *fp, int x) {
    if (fwrite(&x, sizeof(int), 1, fp) != 1) die("cant write binary int");
}
void write_bvec(FILE *fp, const double *a, int m) {
    write_bint(fp, m);
    if (m > 0 && fwrite(a, sizeof(double), m, fp) != (size_t)m) die("cant write bvec");
}
void write_bivec(FILE *fp, const int *a, int m) {
    write_bint(fp, m);
    if (m > 0 && fwrite(a, sizeof(int), m, fp) != (size_t)m) die("cant write bivec");
}
void write_bmat(FILE *fp, const double *a, int m, int n) {
    write_bint(fp, m);
    write_bint(fp, n);
    size_t cnt = (size_t)m * n;
    if (cnt && fwrite(a, sizeof(double), cnt, fp) != cnt) die("cant write bmat");
}

// ------------------------------------------------------------------ foreign formats
// Parallel parse of the tree blocks of a PAML .lnf file (src/freadmat.c:110-154 reads them
// with fscanf: a K=1000 x 10^5-pattern file is gigabytes of text).  The rest of the file is read into
// memory and cut into lines; tree i must be the non-blank line i*(p+1) holding the single integer i+1,
// followed by p pattern lines `j count lnL ...` -- every such block is parsed on its own thread with
// std::from_chars.  Anything else (comments, short files, wrong indices, hex floats) makes it give up
// WITHOUT consuming input, so the serial loop below produces the reference's values and error messages.
static bool read_paml_blocks_parallel(FILE *fp, int m, int n, int p, double *out) {
    if ((long long)m * p < (1 << 15)) return false;
    FileTail file;
    if (!file.load(fp)) return false;
    const size_t bytes = file.bytes;
    const long pos = file.pos;
    const char *base = file.buf.get();
    // offsets of the non-blank lines: every thread indexes a slice that starts at a line start, the
    // slices' lists are then laid end to end
    std::vector<size_t> line;
    {
        const int nt = io_threads();
        std::vector<size_t> cut(nt + 1, bytes);
        cut[0] = 0;
        for (int t = 1; t < nt; t++) {
            size_t c = std::max(cut[t - 1], bytes / nt * t);
            const void *nl = c < bytes ? memchr(base + c, '\n', bytes - c) : nullptr;
            cut[t] = nl ? (size_t)(static_cast<const char *>(nl) - base) + 1 : bytes;
        }
        std::vector<std::vector<size_t>> local(nt);
        auto index = [&](int t) {
            std::vector<size_t> &v = local[t];
            try {
                v.reserve((cut[t + 1] - cut[t]) / 24 + 16);
            } catch (...) {
            }
            for (size_t c = cut[t]; c < cut[t + 1];) {
                const size_t e = (size_t)(static_cast<const char *>(memchr(base + c, '\n', bytes + 1 - c)) - base);  // the sentinel ends the last line
                size_t k = c;
                while (k < e && is_space((unsigned char)base[k])) k++;
                if (k < e) v.push_back(c);
                c = e + 1;
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(index, t);
        index(0);
        for (auto &x : th) x.join();
        size_t total = 0;
        std::vector<size_t> off(nt, 0);
        for (int t = 0; t < nt; t++) {
            off[t] = total;
            total += local[t].size();
        }
        line.resize(total);
        std::vector<std::thread> th2;
        auto place = [&](int t) {
            if (!local[t].empty()) memcpy(line.data() + off[t], local[t].data(), local[t].size() * sizeof(size_t));
        };
        for (int t = 1; t < nt; t++) th2.emplace_back(place, t);
        place(0);
        for (auto &x : th2) x.join();
    }
    const size_t need = (size_t)m * ((size_t)p + 1);
    bool ok = line.size() >= need;
    if (ok) {
        const int nthr = std::max(1, std::min(io_threads(), m));
        std::vector<char> good(nthr, 1);
        auto tok_int = [&](const char *&c, long long *v) {
            while (*c == ' ' || *c == '\t' || *c == '\r') c++;
            const char *b = c;
            while (*c != '\n' && !is_space((unsigned char)*c)) c++;
            auto r = std::from_chars(b, c, *v);
            return b < c && r.ec == std::errc() && r.ptr == c;
        };
        auto work = [&](int t) {
            for (int i = t; i < m; i += nthr) {
                const char *c = base + line[(size_t)i * (p + 1)];
                long long v = 0;
                if (*c == '#' || !tok_int(c, &v) || v != i + 1) { good[t] = 0; return; }
                while (*c == ' ' || *c == '\t' || *c == '\r') c++;
                if (*c != '\n') { good[t] = 0; return; }  // the tree line carries the index only
                int col = 0;
                for (int j = 0; j < p; j++) {
                    c = base + line[(size_t)i * (p + 1) + 1 + j];
                    long long idx = 0, reps = 0;
                    if (*c == '#' || !tok_int(c, &idx) || idx != j + 1 || !tok_int(c, &reps) || reps < 0 || col + reps > n) { good[t] = 0; return; }
                    while (*c == ' ' || *c == '\t' || *c == '\r') c++;
                    const char *b = c;
                    while (*c != '\n' && !is_space((unsigned char)*c)) c++;
                    const char *b2 = (*b == '+') ? b + 1 : b;
                    double x;
                    auto r = std::from_chars(b2, c, x);
                    if (b2 >= c || r.ec != std::errc() || r.ptr != c || (b2 + 1 < c && b2[0] == '0' && (b2[1] == 'x' || b2[1] == 'X'))) { good[t] = 0; return; }
                    for (long long k = 0; k < reps; k++) out[(size_t)i * n + col++] = x;
                }
                if (col != n) { good[t] = 0; return; }
            }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < nthr; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
        for (int t = 0; t < nthr; t++) ok = ok && good[t];
    }
    if (!ok) {
        fseek(fp, pos, SEEK_SET);
        return false;
    }
    // leave the stream behind the last pattern line, as the serial reader does
    size_t end = line[need - 1];
    while (base[end] != '\n') end++;
    fseek(fp, pos + (long)std::min(end + 1, bytes), SEEK_SET);
    return true;
}

std::vector<double> read_mat_paml(FILE *fp, int *mp, int *np) {
    int m = read_int(fp), n = read_int(fp), p = read_int(fp);
    dprint(2, "\nm=%d n=%d p=%d", m, n, p);
    if (m < 0) die("negative tree numbers");
    if (*mp > 0 && *mp != m) die("size of rows mismatch in mat");
    if (*np > 0 && *np != n) die("size of columns mismatch in mat");
    skip_line(fp);
    std::vector<double> a;
    if ((long long)m * n == 0) return a;
    a.resize((size_t)m * n);
    if (debugmode < 2 && p > 0 && read_paml_blocks_parallel(fp, m, n, p, a.data())) {
        *mp = m;
        *np = n;
        return a;
    }
    for (int i = 0; i < m; i++) {
        int t = read_int(fp);
        dprint(2, "\ni=%d", t);
        if (t != i + 1) die("wrong row index");
        int col = 0;
        for (int j = 0; j < p; j++) {
            t = read_int(fp);
            dprint(2, "\n%d", t);
            if (t != j + 1) die("wrong column index");
            int reps = read_int(fp);
            double x = read_double(fp);
            dprint(2, " %d %lg", reps, x);
            skip_line(fp);  // only the first three columns of a pattern line are used
            for (int k = 0; k < reps; k++) {
                if (col >= n) die("too many repeats");
                a[(size_t)i * n + col++] = x;
            }
        }
        if (col != n) die("too few repeats");
    }
    *mp = m;
    *np = n;
    return a;
}

std::vector<double> read_mat_molphy(FILE *fp, int *mp, int *np) {
    int m = read_int(fp), n = read_int(fp);
    skip_line(fp);
    if (*mp > 0 && *mp != m) die("size of rows mismatch in mat");
    if (*np > 0 && *np != n) die("size of columns mismatch in mat");
    std::vector<double> a;
    if ((long long)m * n == 0) return a;
    a.resize((size_t)m * n);
    for (size_t e = 0; e < a.size(); e++) a[e] = read_double(fp);
    *mp = m;
    *np = n;
    return a;
}

std::vector<double> read_mat_puzzle(FILE *fp, int *mp, int *np) {
    int m = read_int(fp), n = read_int(fp);
    dprint(2, "\nm=%d n=%d", m, n);
    skip_line(fp);
    if (*mp > 0 && *mp != m) die("size of rows mismatch in mat");
    if (*np > 0 && *np != n) die("size of columns mismatch in mat");
    std::vector<double> a;
    if ((long long)m * n == 0) return a;
    a.resize((size_t)m * n);
    for (int i = 0; i < m; i++) {
        skip_junk(fp);
        skip_word(fp);  // tree label
        dprint(2, "\ni=%d", i + 1);
        for (int j = 0; j < n; j++) {
            a[(size_t)i * n + j] = read_double(fp);
            dprint(2, "\n%d %lg", j + 1, a[(size_t)i * n + j]);
        }
    }
    *mp = m;
    *np = n;
    return a;
}

// read up to `length` characters of the current line (the newline is consumed only if
// reached); returns the number of characters stored              (src/freadmat.c:156-167)
static int read_line_prefix(FILE *fp, char *buf, int length) {
    int i;
    for (i = 0; i < length; i++) {
        int c = getc(fp);
        if (c == EOF || c == '\n') break;
        buf[i] = (char)c;
    }
    return i;
}
static int read_int_or_zero(FILE *fp) {
    int x = 0;
    skip_junk(fp);
    if (fscanf(fp, "%d", &x) != 1) x = 0;
    return x;
}
static bool line_starts_with(FILE *fp, const char *head) {
    // the reference reads sizeof(head) characters (one more than strlen) and requires that
    // many to be present before comparing the first strlen(head)
    const int want = (int)strlen(head) + 1;
    std::vector<char> buf(want + 1, 0);
    int r = read_line_prefix(fp, buf.data(), want);
    skip_line(fp);
    return r == want && strncmp(head, buf.data(), want - 1) == 0;
}

static std::vector<double> rows_to_matrix(const std::vector<std::vector<double>> &rows, int *mp, int *np) {
    int m = (int)rows.size();
    if (*mp > 0 && *mp != m) die("size of rows mismatch in mat");
    *mp = m;
    std::vector<double> a;
    for (auto &r : rows) a.insert(a.end(), r.begin(), r.end());
    return a;
}

// PAUP "Single-site" blocks                                       (src/freadmat.c:205-233)
static std::vector<double> read_mat_paup_blocks(FILE *fp, int *mp, int *np) {
    std::vector<std::vector<double>> rows;
    for (;;) {
        bool found = false;
        while (!feof(fp) && !ferror(fp)) {
            skip_junk(fp);
            if (line_starts_with(fp, "Single-site")) {
                found = true;
                break;
            }
        }
        if (!found) break;
        std::vector<double> v;
        for (; !feof(fp) && !ferror(fp); skip_junk(fp)) {
            int t = read_int_or_zero(fp);
            if (t != (int)v.size() + 1) break;
            v.push_back(read_double(fp));
        }
        if (*np > 0 && *np != (int)v.size()) die("size of columns mismatch in mat");
        *np = (int)v.size();
        rows.push_back(v);
    }
    return rows_to_matrix(rows, mp, np);
}

// PAUP tab-separated table, values negated, gaps dropped           (src/freadmat.c:235-279)
static std::vector<double> read_mat_paup_table(FILE *fp, int *mp, int *np) {
    std::vector<std::vector<double>> rows;
    std::vector<double> v;
    while (!feof(fp) && !ferror(fp)) {
        int c = getc(fp);
        if (c == '#') {
            skip_line(fp);
            continue;
        }
        if (c == '\t') {
            (void)read_int_or_zero(fp);
            skip_junk(fp);
            double t;
            int d = getc(fp);
            if (d == '-') t = -1.0;  // a gap column
            else {
                ungetc(d, fp);
                t = read_double(fp);
            }
            if (t > 0) v.push_back(-t);
            skip_line(fp);
        } else {
            skip_line(fp);
            if (!v.empty()) {
                if (*np > 0 && *np != (int)v.size()) die("size of columns mismatch in mat");
                *np = (int)v.size();
                rows.push_back(v);
                v.clear();
            }
        }
    }
    return rows_to_matrix(rows, mp, np);
}

std::vector<double> read_mat_paup(FILE *fp, int *mp, int *np) {
    const char *h2 = "Tree\t-lnL\tSite", *h3 = "Tree\tLength\tCh";
    const int want2 = (int)strlen(h2) + 1, want3 = (int)strlen(h3) + 1;
    char buf[32] = {0};
    int r = read_line_prefix(fp, buf, want2);
    skip_line(fp);
    if ((r == want2 && strncmp(h2, buf, want2 - 1) == 0) || (r == want3 && strncmp(h3, buf, want3 - 1) == 0))
        return read_mat_paup_table(fp, mp, np);
    return read_mat_paup_blocks(fp, mp, np);
}

// PhyML: blocks headed "Site   P(D|M)", value log(P)               (src/freadmat.c:304-360)
std::vector<double> read_mat_phyml(FILE *fp, int *mp, int *np) {
    std::vector<std::vector<double>> rows;
    for (;;) {
        bool found = false;
        while (!feof(fp) && !ferror(fp)) {
            skip_junk(fp);
            if (line_starts_with(fp, "Site   P(D|M)")) {
                found = true;
                break;
            }
        }
        if (!found) break;
        std::vector<double> v;
        while (!feof(fp) && !ferror(fp)) {
            int t = read_int_or_zero(fp);
            if (t == 0) break;
            double x = read_double(fp);
            skip_line(fp);
            v.push_back(log(x));
        }
        if (!v.empty()) {
            if (*np > 0 && *np != (int)v.size()) die("size of columns mismatch in mat");
            *np = (int)v.size();
        }
        rows.push_back(v);
    }
    return rows_to_matrix(rows, mp, np);
}

// ------------------------------------------------------------------ sort
void sort_with_order(std::vector<double> &x, std::vector<int> *order) {
    // src/misc.c:657-700: same partition scheme as psort_order, on values
    const int dim = (int)x.size();
    std::vector<int> ord(dim);
    for (int i = 0; i < dim; i++) ord[i] = i;
    if (dim > 1) {
        int stack[64], sp = 0, l = 0, r = dim - 1;
        for (;;) {
            while (r > l) {
                int k = (r + l) / 2;
                const double v = x[k];
                int i = l - 1, j = r + 1;
                for (;;) {
                    while (i < r && x[++i] <= v) {}
                    while (j > l && x[--j] >= v) {}
                    if (i > j) break;
                    if (i == r) j = k; else if (j == l) i = k;
                    std::swap(x[i], x[j]);
                    std::swap(ord[i], ord[j]);
                    if (i == r) { j = r - 1; break; } else if (j == l) { i = l + 1; break; }
                }
                if (j - l > r - i) { stack[sp++] = l; stack[sp++] = j; l = i; }
                else { stack[sp++] = i; stack[sp++] = r; r = j; }
            }
            if (sp == 0) break;
            r = stack[--sp];
            l = stack[--sp];
        }
    }
    if (order) *order = ord;
}

}  // namespace refio
