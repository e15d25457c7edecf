// pb_format.cpp -- text of one window exactly as the reference's printers write it:
//   print_nucdiv  pop_nucdiv.cpp:258-289     print_sfs     pop_sfs.cpp:293-317
//   print_ld      pop_ld.cpp:650-712         print_diverge pop_diverge.cpp:496-574
//   print_haplo   pop_haplo.cpp:365-442      print_popbam_snp / print_sweep / print_ms  pop_snp.cpp:224-303
// The reference streams with std::fixed << std::setprecision(5) and prints "NA" through std::setw(7);
// "%.5f" and "%7s" produce the same bytes.  Host-only formatting of results the kernels computed.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/popbam_b200.h"

extern "C" const pb_params *pb_ctx_params(const pb_ctx *c);

namespace {

struct Out {
    std::string s;
    void f(const char *fmt, ...) __attribute__((format(printf, 2, 3))) {
        char tmp[640];
        va_list ap;
        va_start(ap, fmt);
        int k = vsnprintf(tmp, sizeof tmp, fmt, ap);
        va_end(ap);
        if (k > 0) s.append(tmp, (size_t)std::min<int>(k, (int)sizeof tmp - 1));
    }
    // "\tname[label]:\t<value or NA>"
    void stat(const char *name, const char *label, bool ok, double v) {
        if (ok) f("\t%s[%s]:\t%.5f", name, label, v);
        else f("\t%s[%s]:\t%7s", name, label, "NA");
    }
};

// bam_nt16_table followed by bam_nt16_rev_table (popbam.cpp:13-31, bam.h): the letter the reference
// prints for a reference byte / consensus letter
char nt16_roundtrip(int c) {
    static const char rev[17] = "=ACMGRSVTWYHKDBN";
    int code;
    switch (c) {
    case '=': code = 0; break;
    case 'A': case 'a': case '0': code = 1; break;
    case 'C': case 'c': case '1': code = 2; break;
    case 'M': case 'm': code = 3; break;
    case 'G': case 'g': case '2': code = 4; break;
    case 'R': case 'r': code = 5; break;
    case 'S': case 's': code = 6; break;
    case 'V': case 'v': code = 7; break;
    case 'T': case 't': case '3': code = 8; break;
    case 'W': case 'w': code = 9; break;
    case 'Y': case 'y': code = 10; break;
    case 'H': case 'h': code = 11; break;
    case 'K': case 'k': code = 12; break;
    case 'D': case 'd': code = 13; break;
    case 'B': case 'b': code = 14; break;
    default: code = 15; break;
    }
    return rev[code];
}
const char kIupacLetters[17] = "AMRWNCSYNNGKNNNT";   // popbam.cpp iupac[]

inline int popc(uint64_t x) { return __builtin_popcountll(x); }

}  // namespace

extern "C" int64_t pb_format_window(const pb_ctx *ctx, const pb_region_result *r, int32_t w, uint32_t an, const pb_print_opts *o,
                                    char *buf, int64_t cap) {
    const pb_params *p = pb_ctx_params(ctx);
    if (!p || !r || !o || w < 0 || w >= r->n_windows) return PB_ERR_ARG;
    if (!(r->analyses & an)) return PB_ERR_ARG;
    Out out;
    const int P = r->n_pops, n = r->n_samples;
    const int ns = r->num_sites[w];
    const int64_t so = r->seg_off[w];
    const int S = r->segsites[w];
    const size_t wp = (size_t)w * P;
    if (an == PB_AN_SNP) {
        if (!r->seg_cb) return PB_ERR_ARG;
        if (o->snp_output == 0) {                    // print_popbam_snp
            for (int i = 0; i < S; ++i) {
                out.f("%s\t%u\t%c", o->chrom, r->seg_pos[so + i] + 1, nt16_roundtrip(r->seg_ref[so + i]));
                for (int j = 0; j < n; ++j) {
                    const uint64_t cb = r->seg_cb[(size_t)(so + i) * n + j];
                    const unsigned g = (unsigned)((cb >> 8) & 0xff);
                    // g >= 16 only arises from the reference's revert-to-reference arithmetic,
                    // where the reference itself reads iupac[] out of bounds: that letter is undefined
                    const char base = g < 16 ? nt16_roundtrip(kIupacLetters[g]) : '?';
                    out.f("\t%c\t%u\t%u\t%u", base, (unsigned)((cb >> 32) & 0xffff), (unsigned)((cb >> 48) & 0xffff),
                          (unsigned)((cb >> 16) & 0xffff));
                }
                out.f("\n");
            }
        } else if (o->snp_output == 1) {             // print_sweep
            for (int i = 0; i < S; ++i) {
                out.f("%s\t%u", o->chrom, r->seg_pos[so + i] + 1);
                const uint64_t T = r->seg_type[so + i];
                for (int j = 0; j < P; ++j) {
                    const int pn = popc(p->pop_mask[j]);
                    unsigned f = (unsigned)popc(T & p->pop_mask[j]);
                    if ((p->flags & PB_FLAG_OUTGROUP) && (T >> p->outidx & 1)) f = (unsigned)(pn - (int)f) & 0xffff;
                    out.f("\t%u\t%d", f, pn);
                }
                out.f("\n");
            }
        } else {                                     // print_ms
            out.f("//\nsegsites: %d\npositions: ", S);
            for (int i = 0; i < S; ++i)
                out.f("%.8g ", (double)(r->seg_pos[so + i] - (unsigned)r->win_beg[w]) / (r->win_end[w] - r->win_beg[w]));
            out.f("\n");
            for (int i = 0; i < n; ++i) {
                for (int j = 0; j < S; ++j) {
                    const uint64_t T = r->seg_type[so + j];
                    int bit = (int)(T >> i & 1);
                    if ((p->flags & PB_FLAG_OUTGROUP) && (T >> p->outidx & 1)) bit = !bit;
                    out.s.push_back(bit ? '1' : '0');
                }
                out.f("\n");
            }
            out.f("\n");
        }
    } else {
        out.f("%s\t%d\t%d\t%d", o->chrom, r->win_beg[w] + 1, r->win_end[w] + 1, ns);
        const bool ok = ns >= o->min_sites;
        char nm[600];
        switch (an) {
        case PB_AN_NUCDIV:
            for (int i = 0; i < P; ++i) out.stat("pi", o->pop_names[i], ok, r->piw[wp + i] / ns);
            for (int i = 0; i < P - 1; ++i)
                for (int j = i + 1; j < P; ++j) {
                    snprintf(nm, sizeof nm, "%s-%s", o->pop_names[i], o->pop_names[j]);
                    out.stat("dxy", nm, ok, r->pib[wp * P + i * P + (j - (i + 1))] / ns);
                }
            break;
        case PB_AN_SFS:
            for (int i = 0; i < P; ++i) {
                out.stat("D", o->pop_names[i], !std::isnan(r->td[wp + i]), r->td[wp + i]);
                out.stat("H", o->pop_names[i], !std::isnan(r->fwh[wp + i]), r->fwh[wp + i]);
            }
            break;
        case PB_AN_LD_ZNS: case PB_AN_LD_OMEGA: case PB_AN_LD_WALL:
            for (int i = 0; i < P; ++i) {
                const int nsnp = an == PB_AN_LD_WALL ? r->wall_num_snps[wp + i] : r->ld_num_snps[wp + i];
                out.f("\tS[%s]:\t%d", o->pop_names[i], nsnp);
                const bool good = nsnp >= o->min_snps;
                if (an == PB_AN_LD_ZNS) out.stat("Zns", o->pop_names[i], good, r->zns[wp + i]);
                else if (an == PB_AN_LD_OMEGA) out.stat("omax", o->pop_names[i], good, r->omegamax[wp + i]);
                else { out.stat("B", o->pop_names[i], good, r->wallb[wp + i]); out.stat("Q", o->pop_names[i], good, r->wallq[wp + i]); }
            }
            break;
        case PB_AN_DIVERGE_IND:
            for (int i = 0; i < n; ++i) {
                double d = (double)r->ind_div[(size_t)w * n + i] / ns;
                if (o->jc) d = -0.75 * std::log(1.0 - d * (4.0 / 3.0));
                out.stat("d", o->sample_names[i], ok, d);
            }
            break;
        case PB_AN_DIVERGE_POP:
            for (int i = 0; i < P; ++i) {
                const char *pn = o->pop_names[i];
                if (ok) {
                    const int fx = r->pop_div[wp + i], sg = r->div_num_snps[wp + i];
                    double d = (p->flags & PB_FLAG_SUBSTITUTE) ? (double)fx / ns : (double)(fx + sg) / ns;
                    if (o->jc) d = -0.75 * std::log(1.0 - d * (4.0 / 3.0));
                    out.f("\tFixed[%s]:\t%d\tSeg[%s]:\t%d\td[%s]:\t%.5f", pn, fx, pn, sg, pn, d);
                } else out.f("\tFixed[%s]:\t%7s\tSeg[%s]:\t%7s\td[%s]:\t%7s", pn, "NA", pn, "NA", pn, "NA");
            }
            break;
        case PB_AN_HAPLO_K:
            for (int i = 0; i < P; ++i) {
                const char *pn = o->pop_names[i];
                if (ok) out.f("\tK[%s]:\t%d\tKdiv[%s]:\t%.5f", pn, r->nhaps[wp + i], pn, 1.0 - r->hdiv[wp + i]);
                else out.f("\tK[%s]:\t%7s\tKdiv[%s]:\t%7s", pn, "NA", pn, "NA");
            }
            break;
        case PB_AN_HAPLO_EHHS:
            for (int i = 0; i < P; ++i) out.stat("EHHS", o->pop_names[i], ok && !std::isnan(r->ehhs[wp + i]), r->ehhs[wp + i]);
            break;
        case PB_AN_HAPLO_DXY:
            for (int i = 0; i < P; ++i) out.stat("pi", o->pop_names[i], ok, r->piw[wp + i]);
            for (int i = 0; i < P - 1; ++i)
                for (int j = i + 1; j < P; ++j) {
                    snprintf(nm, sizeof nm, "%s-%s", o->pop_names[i], o->pop_names[j]);
                    const size_t x = wp * P + i * P + (j - (i + 1));
                    if (ok) out.f("\tdxy[%s]:\t%.5f\tmin[%s]:\t%u", nm, r->pib[x], nm, (unsigned)r->min_dxy[x]);
                    else out.f("\tdxy[%s]:\t%7s\tmin[%s]:\t%7s", nm, "NA", nm, "NA");
                }
            break;
        default: return PB_ERR_ARG;
        }
        out.f("\n");
    }
    const int64_t len = (int64_t)out.s.size();
    if (buf && cap > 0) {
        const int64_t k = std::min<int64_t>(len, cap - 1);
        memcpy(buf, out.s.data(), (size_t)k);
        buf[k] = 0;
    }
    return len;
}
