// pb_format.cpp -- text of one window exactly as the reference's printers write it:
//   print_nucdiv  pop_nucdiv.cpp:258-289     print_sfs     pop_sfs.cpp:293-317
//   print_ld      pop_ld.cpp:650-712         print_diverge pop_diverge.cpp:496-574
//   print_haplo   pop_haplo.cpp:365-442      print_popbam_snp / print_sweep / print_ms  pop_snp.cpp:224-303
//   make_nj / join_tree / print_tree  pop_tree.cpp:208-470  (neighbour joining on the window's difference matrix)
// The reference streams with std::fixed << std::setprecision(5) and prints "NA" through std::setw(7);
// "%.5f" and "%7s" produce the same bytes.  Host-only formatting of results the kernels computed.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <string>
#include <vector>
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

// Neighbour joining as the tree subcommand does it (calc_dist_matrix, join_tree, print_tree: pop_tree.cpp:496-515,
// 254-429, 439-470), on flat arrays: taxon t (0 = the reference sequence) is end point t; interior node k has the three
// end points T + 3k + {0, 1, 2}, of which 1 and 2 receive the two clusters it joins and 0 is joined later.  `mate[e]` is
// the end point at the other end of e's branch, `len[e]` the branch length (stored at both ends).  The arithmetic keeps
// the reference's order of operations: ties in the minimisation and the printed lengths depend on it.
struct NjTree {
    int T;
    std::vector<int> mate;
    std::vector<double> len;
    const pb_print_opts *o;
    void link(int a, int b) { mate[a] = b; mate[b] = a; }
    bool tip(int e) const { return e < T; }
    int next(int e) const { const int k = (e - T) / 3; return T + 3 * k + ((e - T) % 3 + 1) % 3; }
    void print(Out &out, int e, int start) const {
        if (tip(e)) out.s += e == 0 ? o->ref_name : o->sample_names[e - 1];
        else {
            out.s += '(';
            print(out, mate[next(e)], start);
            out.s += ',';
            print(out, mate[next(next(e))], start);
            if (e == start) { out.s += ','; print(out, mate[e], start); }
            out.s += ')';
        }
        if (e == start) out.s += ";\n";
        else if (len[e] < 0) out.s += ":0.00000";
        else out.f(":%.5f", len[e]);
    }
};

void nj_newick(Out &out, const uint16_t *diff, int T, int num_sites, const pb_print_opts *o) {
    std::vector<double> x((size_t)T * T, 0.0), av((size_t)T, 0.0), R((size_t)T);
    auto X = [&](int a, int b) -> double & { return x[(size_t)a * T + b]; };
    for (int i = 0; i < T - 1; ++i)
        for (int j = i + 1; j < T; ++j) {
            double d = (double)diff[i * T + j] / num_sites;                 // p-distance
            if (o->jc) d = -0.75 * log(1.0 - (4.0 * d / 3.0));              // Jukes-Cantor
            X(i, j) = d; X(j, i) = d;
        }
    NjTree t;
    t.T = T; t.o = o;
    t.mate.assign((size_t)T + 3 * (size_t)(T - 2), -1);
    t.len.assign(t.mate.size(), 0.0);
    std::vector<int> cl((size_t)T);          // live clusters: the end point that represents each, -1 when merged away
    for (int i = 0; i < T; ++i) cl[i] = i;
    for (int i = 0; i < T - 1; ++i)
        for (int j = i + 1; j < T; ++j) { const double da = (X(i, j) + X(j, i)) / 2.0; X(i, j) = da; X(j, i) = da; }
    double fotu2 = T - 2.0, total = 0.0;
    int node = 0, mi = 0, mj = 0;
    for (int cycle = 1; cycle <= T - 3; ++cycle, ++node) {
        for (int j = 1; j < T; ++j)
            for (int i = 0; i < j; ++i) X(j, i) = X(i, j);
        double tmin = DBL_MAX;
        for (int i = 0; i < T; ++i) R[i] = 0.0;
        for (int j = 1; j < T; ++j) {
            if (cl[j] < 0) continue;
            for (int i = 0; i < j; ++i)
                if (cl[i] >= 0) { R[i] += X(i, j); R[j] += X(i, j); }
        }
        for (int j = 1; j < T; ++j) {
            if (cl[j] < 0) continue;
            for (int i = 0; i < j; ++i) {
                if (cl[i] >= 0) total = fotu2 * X(i, j) - R[i] - R[j];
                if (total < tmin) { tmin = total; mi = i; mj = j; }          // the reference compares a stale total for dead i too
            }
        }
        double dio = 0.0, djo = 0.0;
        for (int i = 0; i < T; ++i) { dio += X(i, mi); djo += X(i, mj); }
        const double dmin = X(mi, mj);
        dio = (dio - dmin) / fotu2;
        djo = (djo - dmin) / fotu2;
        double bi = (dmin + dio - djo) * 0.5, bj = dmin - bi;
        bi -= av[mi]; bj -= av[mj];
        const int e0 = T + 3 * node;
        t.link(e0 + 1, cl[mi]); t.link(e0 + 2, cl[mj]);
        t.len[cl[mi]] = bi; t.len[e0 + 1] = bi;
        t.len[cl[mj]] = bj; t.len[e0 + 2] = bj;
        cl[mi] = e0; cl[mj] = -1;
        av[mi] = dmin * 0.5;
        fotu2 -= 1.0;
        for (int j = 0; j < T; ++j)
            if (cl[j] >= 0) {
                const double da = (X(mi, j) + X(mj, j)) * 0.5;
                if (mi < j) X(mi, j) = da;
                if (mi > j) X(j, mi) = da;
            }
        for (int j = 0; j < T; ++j) { X(mj, j) = 0.0; X(j, mj) = 0.0; }
    }
    int el[3], k = 0;
    for (int i = 0; i < T && k < 3; ++i) if (cl[i] >= 0) el[k++] = i;
    double bi = (X(el[0], el[1]) + X(el[0], el[2]) - X(el[1], el[2])) * 0.5;
    double bj = X(el[0], el[1]) - bi, bk = X(el[0], el[2]) - bi;
    bi -= av[el[0]]; bj -= av[el[1]]; bk -= av[el[2]];
    const int e0 = T + 3 * node;
    t.link(e0, cl[el[0]]); t.link(e0 + 1, cl[el[1]]); t.link(e0 + 2, cl[el[2]]);
    t.len[cl[el[0]]] = bi; t.len[e0] = bi;
    t.len[cl[el[1]]] = bj; t.len[e0 + 1] = bj;
    t.len[cl[el[2]]] = bk; t.len[e0 + 2] = bk;
    t.print(out, t.mate[0], t.mate[0]);      // make_nj starts at the node the reference taxon hangs on
}

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
        if (an == PB_AN_TREE) {      // make_nj (pop_tree.cpp:208-252): the tree ends the row itself
            if (!r->tree_diff || !o->ref_name || n < 2) return PB_ERR_ARG;      // three taxa at least
            if (!ok || S < 1) out.f("\tNA\n");
            else { out.f("\t"); nj_newick(out, r->tree_diff + (size_t)w * (n + 1) * (n + 1), n + 1, ns, o); }
        } else {
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
    }
    const int64_t len = (int64_t)out.s.size();
    if (buf && cap > 0) {
        const int64_t k = std::min<int64_t>(len, cap - 1);
        memcpy(buf, out.s.data(), (size_t)k);
        buf[k] = 0;
    }
    return len;
}
