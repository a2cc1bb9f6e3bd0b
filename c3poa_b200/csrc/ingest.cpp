// ingest.cpp -- FASTQ/FASTA reader + length filter + batcher.
//
// Replaces the reference's use of mappy for I/O (mm.fastx_read, /root/reference/C3POa.py:201-206,239-254):
// records shorter than --lencutoff are counted and skipped, the rest land back to back in caller-owned
// (ideally pinned) buffers in exactly the layout c3_stage() consumes, plus the per-read Phred sum the
// header needs (C3POa.py:168).  Host code only; plain or gzip input through zlib.
#include "../../include/c3poa_gpu.h"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

struct c3_fastq {
    gzFile f = nullptr;             // gzip input
    FILE *fp = nullptr;             // plain input: own buffered line reader (memchr), ~10x faster than gzgets
    std::string buf; size_t bpos = 0, blen = 0;
    std::string line, name, seq, qual;
    bool have_pending = false;      // a parsed record that did not fit the previous batch
    bool eof = false;
    std::string carry;              // a header line read ahead (FASTA)
};

static bool read_line_plain(c3_fastq *q, std::string &out)
{
    out.clear();
    for (;;) {
        if (q->bpos == q->blen) {
            q->blen = fread(&q->buf[0], 1, q->buf.size(), q->fp);
            q->bpos = 0;
            if (q->blen == 0) return !out.empty();
        }
        const char *p = &q->buf[q->bpos];
        const char *nl = (const char *)memchr(p, '\n', q->blen - q->bpos);
        if (nl) {
            size_t n = (size_t)(nl - p);
            out.append(p, n);
            q->bpos += n + 1;
            if (!out.empty() && out.back() == '\r') out.pop_back();
            return true;
        }
        out.append(p, q->blen - q->bpos);
        q->bpos = q->blen;
    }
}

static bool read_line(c3_fastq *q, std::string &out)
{
    if (q->fp) return read_line_plain(q, out);
    out.clear();
    char buf[65536];
    for (;;) {
        if (!gzgets(q->f, buf, (int)sizeof(buf))) return !out.empty();
        size_t n = strlen(buf);
        if (n && buf[n - 1] == '\n') {
            --n;
            if (n && buf[n - 1] == '\r') --n;
            out.append(buf, n);
            return true;
        }
        out.append(buf, n);
    }
}

static void set_name(c3_fastq *q, const std::string &hdr)
{
    size_t e = 1;
    while (e < hdr.size() && hdr[e] != ' ' && hdr[e] != '\t') ++e;
    q->name.assign(hdr, 1, e - 1);
}

// parse the next record into q->name/seq/qual; returns false at EOF
static bool next_record(c3_fastq *q)
{
    std::string &ln = q->line;
    for (;;) {
        if (!q->carry.empty()) { ln.swap(q->carry); q->carry.clear(); }
        else if (!read_line(q, ln)) return false;
        if (ln.empty()) continue;
        if (ln[0] == '@') {
            set_name(q, ln);
            if (!read_line(q, q->seq)) return false;
            // multi-line records (sequence and quality wrapped over several lines): sequence lines run up to the '+'
            // line, quality lines until they cover the sequence (a quality line may itself start with '@' or '+')
            std::string nx;
            for (;;) {
                if (!read_line(q, nx)) return false;
                if (!nx.empty() && nx[0] == '+') break;
                q->seq += nx;
            }
            if (!read_line(q, q->qual)) q->qual.clear();
            while (q->qual.size() < q->seq.size() && read_line(q, nx)) q->qual += nx;
            return true;
        }
        if (ln[0] == '>') {
            set_name(q, ln);
            q->seq.clear(); q->qual.clear();
            std::string nx;
            while (read_line(q, nx)) {
                if (!nx.empty() && nx[0] == '>') { q->carry.swap(nx); break; }
                q->seq += nx;
            }
            return true;
        }
    }
}

extern "C" int c3_fastq_open(const char *path, c3_fastq **out)
{
    if (!path || !out) return -1;
    *out = nullptr;
    FILE *fp = fopen(path, "rb");
    if (!fp) return -2;
    unsigned char magic[2] = {0, 0};
    const size_t got = fread(magic, 1, 2, fp);
    c3_fastq *q = new c3_fastq();
    if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
        fclose(fp);
        q->f = gzopen(path, "rb");
        if (!q->f) { delete q; return -2; }
        gzbuffer(q->f, 1 << 20);
    } else {
        rewind(fp);
        q->fp = fp;
        q->buf.resize(4 << 20);
    }
    *out = q;
    return 0;
}

extern "C" void c3_fastq_close(c3_fastq *q)
{
    if (!q) return;
    if (q->f) gzclose(q->f);
    if (q->fp) fclose(q->fp);
    delete q;
}

extern "C" int c3_fastq_next(c3_fastq *q, int32_t max_reads, int64_t max_bases, int32_t min_len,
                             char *seq, char *qual, int64_t *off, char *names, int64_t names_cap,
                             int64_t *name_off, int64_t *qual_sum, int64_t *n_short)
{
    if (!q || !seq || !off || !names || !name_off || max_reads <= 0 || max_bases <= 0) return -1;
    int n = 0;
    int64_t bases = 0, nb = 0;
    off[0] = 0; name_off[0] = 0;
    while (n < max_reads) {
        if (!q->have_pending) {
            if (q->eof || !next_record(q)) { q->eof = true; break; }
            if ((int64_t)q->seq.size() < (int64_t)min_len) { if (n_short) ++*n_short; continue; }
        }
        const int64_t L = (int64_t)q->seq.size(), NL = (int64_t)q->name.size() + 1;
        if (bases + L > max_bases || nb + NL > names_cap) {
            if (n == 0) return -3;                 // a single record larger than the buffers
            q->have_pending = true;
            break;
        }
        q->have_pending = false;
        memcpy(seq + bases, q->seq.data(), (size_t)L);
        long long qs = 0;
        if (qual) {
            const size_t QL = q->qual.size();
            if ((int64_t)QL >= L) {
                memcpy(qual + bases, q->qual.data(), (size_t)L);
                const unsigned char *p = (const unsigned char *)q->qual.data();
                unsigned long long acc = 0;
                for (int64_t i = 0; i < L; ++i) acc += p[i];
                qs = (long long)acc - 33ll * L;
            } else {
                for (int64_t i = 0; i < L; ++i) {
                    const char c = (size_t)i < QL ? q->qual[(size_t)i] : 'I';  // FASTA: constant quality
                    qual[bases + i] = c;
                    qs += (unsigned char)c - 33;
                }
            }
        } else {
            for (char c : q->qual) qs += (unsigned char)c - 33;
        }
        if (qual_sum) qual_sum[n] = qs;
        memcpy(names + nb, q->name.c_str(), (size_t)NL);
        bases += L; nb += NL; ++n;
        off[n] = bases; name_off[n] = nb;
    }
    return n;
}

// ---------------------------------------------------------------------------
// Output side: the consensus FASTA and subread FASTQ records of a batch, formatted
// exactly as the reference writes them --
//   >{name}_{avg_qual}_{len}_{repeats}_{cons_len}\n{cons}\n                      (C3POa.py:167-173)
//   @{name}_{k}\n{subread}\n+\n{qual}\n   k = 1..repeats; dangling ends k = 0 and k = repeats + 1
//                                                                                (bin/determine_consensus.py:57-77)
// -- for the reads of one output group (splint directory).  Python formatting of 50 000 reads per batch was the
// slowest stage of the driver by 10x; this is one pass of memcpy.
// ---------------------------------------------------------------------------
static char *put_int(char *p, long long v)
{
    char tmp[24]; int n = 0;
    if (v < 0) { *p++ = '-'; v = -v; }
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

// str(round(x, 2)) of Python for x >= 0: two decimals correctly rounded, trailing zeros dropped, at least one kept
static char *put_avg_qual(char *p, long long qsum, long long len)
{
    char tmp[64];
    int n = snprintf(tmp, sizeof(tmp), "%.2f", (double)qsum / (double)len);
    while (n > 0 && tmp[n - 1] == '0' && n >= 2 && tmp[n - 2] != '.') --n;
    memcpy(p, tmp, (size_t)n);
    return p + n;
}

extern "C" int c3_format_batch(int32_t n_reads, const char *names, const int64_t *name_off,
                               const char *seq, const char *qual, const int64_t *off, const int64_t *qual_sum,
                               const c3_read_result *res, const int32_t *sub_bounds, const int32_t *dang_bounds,
                               int32_t max_peaks, const char *cons, int32_t cons_cap,
                               const int32_t *group, int32_t which_group,
                               char *out_cons, int64_t out_cons_cap, int64_t *out_cons_len,
                               char *out_sub, int64_t out_sub_cap, int64_t *out_sub_len, int64_t *stats)
{
    if (n_reads < 0 || !names || !name_off || !seq || !off || !qual_sum || !res || !sub_bounds || !dang_bounds || !cons ||
        !out_cons || !out_cons_len || !out_sub || !out_sub_len)
        return -1;
    char *pc = out_cons, *ps = out_sub;
    long long n_cons = 0, n_nopeak = 0, n_left = 0, n_err = 0;
    for (int32_t i = 0; i < n_reads; ++i) {
        if (group && group[i] != which_group) continue;
        const c3_read_result &r = res[i];
        if (r.status == 1) { ++n_nopeak; continue; }
        if (r.status < 0) { ++n_err; continue; }
        if (r.status != 0) { ++n_left; continue; }                // 2-repeat / 0-repeat paths: the caller's
        const char *nm = names + name_off[i];
        const size_t nl = strlen(nm);
        const int64_t a0 = off[i], len = off[i + 1] - off[i];
        const int ns = r.n_sub, nd = r.n_dang, cl = r.cons_len;
        if (ns < 0 || ns > max_peaks || nd < 0 || nd > 2 || cl < 0 || cl > cons_cap) return -3;
        if (pc + nl + 96 + cl > out_cons + out_cons_cap) return -2;
        *pc++ = '>'; memcpy(pc, nm, nl); pc += nl; *pc++ = '_';
        pc = put_avg_qual(pc, qual_sum[i], len); *pc++ = '_';
        pc = put_int(pc, len); *pc++ = '_';
        pc = put_int(pc, ns); *pc++ = '_';
        pc = put_int(pc, cl); *pc++ = '\n';
        memcpy(pc, cons + (int64_t)i * cons_cap, (size_t)cl); pc += cl; *pc++ = '\n';
        const int32_t *sb = sub_bounds + (int64_t)i * max_peaks * 2, *db = dang_bounds + (int64_t)i * 4;
        for (int k = 0; k < ns + nd; ++k) {
            const int32_t a = k < ns ? sb[2 * k] : db[2 * (k - ns)], b = k < ns ? sb[2 * k + 1] : db[2 * (k - ns) + 1];
            const int tag = k < ns ? k + 1 : (k == ns ? 0 : ns + 1);
            if (a < 0 || b < a || b > len) return -3;
            const int64_t sl = b - a;
            if (ps + nl + 32 + 2 * sl > out_sub + out_sub_cap) return -2;
            *ps++ = '@'; memcpy(ps, nm, nl); ps += nl; *ps++ = '_';
            ps = put_int(ps, tag); *ps++ = '\n';
            memcpy(ps, seq + a0 + a, (size_t)sl); ps += sl;
            *ps++ = '\n'; *ps++ = '+'; *ps++ = '\n';
            if (qual) { memcpy(ps, qual + a0 + a, (size_t)sl); ps += sl; }
            *ps++ = '\n';
        }
        ++n_cons;
    }
    *out_cons_len = pc - out_cons; *out_sub_len = ps - out_sub;
    if (stats) { stats[0] = n_cons; stats[1] = n_nopeak; stats[2] = n_left; stats[3] = n_err; }
    return 0;
}
