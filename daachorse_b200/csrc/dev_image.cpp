// Builds the device scan image from a validated host automaton.
//
// Results of a scan depend only on the automaton's transition/output structure, never on
// where the builder placed a state, so the image is free to carry
// derived fields next to the reference's BASE / CHECK / FAIL / OUTPUT_POS values.
#include "dev_image.h"

namespace dach {

namespace {

// Every failure chain must end at ROOT (or at DEAD where DEAD is a terminal).  O(n).
bool fail_chains_terminate(const dach_pma* p, bool dead_is_terminal) {
    const size_t n = p->slots();
    std::vector<uint8_t> color(n, 0);  // 0 new, 1 on the current path, 2 proven
    std::vector<uint32_t> path;
    for (size_t start = 0; start < n; ++start) {
        if (color[start]) continue;
        path.clear();
        uint32_t s = uint32_t(start);
        for (;;) {
            if (s == kRoot || (dead_is_terminal && s == kDead) || color[s] == 2) break;
            if (color[s] == 1) return false;  // cycle
            color[s] = 1;
            path.push_back(s);
            s = p->fail[s];
        }
        for (uint32_t v : path) color[v] = 2;
    }
    return true;
}

}  // namespace

int build_image(const dach_pma* p, HostImage* img) {
    const size_t n = p->slots();
    const bool lm = is_leftmost(p->match_kind);
    img->charwise = p->charwise;
    img->match_kind = p->match_kind;
    img->n_slots = uint32_t(n);
    img->root_opos = n ? p->state_output_pos(kRoot) : 0;
    img->max_pattern_len = 0;
    for (const OutputRec& o : p->outputs)
        if (o.length > img->max_pattern_len) img->max_pattern_len = o.length;

    if (n >= 0x80000000ull) {
        set_error("automaton too large for the device image (2^31 slots)");
        return DACH_AUTOMATON_SCALE;
    }
    if (!fail_chains_terminate(p, p->charwise || lm)) {
        set_error("invalid automaton: a failure chain never reaches the root");
        return DACH_INVALID_AUTOMATON;
    }

    img->rec.resize(n * 4);
    if (!p->charwise) {
        auto skip_leaves = [&](uint32_t f) {  // failure target with child-less states skipped
            while (f != kRoot && !(lm && f == kDead) && p->base[f] == 0) f = p->fail[f];
            return f;
        };
        for (size_t s = 0; s < n; ++s) {
            const uint32_t f = skip_leaves(p->fail[s]);
            const bool terminal = f == kRoot || (lm && f == kDead);
            uint32_t* r = &img->rec[s * 4];
            r[0] = p->base[s];
            r[1] = f;
            r[2] = terminal ? 0 : p->base[f];
            // bit 31: the failure state's own (leaf-skipped) failure target is ROOT
            if (!terminal && !lm && skip_leaves(p->fail[f]) == kRoot) r[2] |= 0x80000000u;
            r[3] = p->opos_ch[s];
        }
        // dense ROOT row; also valid for the leftmost automaton, whose ROOT transition is
        // "child or stay" (src/bytewise.rs:1102-1117)
        img->root_table.assign(256, kRoot);
        if (n && p->base[kRoot] != 0) {
            const uint32_t b = p->base[kRoot];
            for (uint32_t c = 0; c < 256; ++c) {
                const uint32_t ci = b ^ c;
                if (ci < n && (p->opos_ch[ci] & 0xff) == c) img->root_table[c] = ci;
            }
        }
        // compact image for the lane-machine kernels (Standard automata of at most 2^24 slots)
        img->root_base = n ? p->base[kRoot] : 0;
        if (n <= (size_t(1) << 24) && p->outputs.size() < (size_t(1) << 24)) {
            img->crec.resize(n * 4);
            img->opos_tab.resize(n);
            for (size_t s = 0; s < n; ++s) {
                const uint32_t* w = &img->rec[s * 4];
                uint32_t sig = 0;
                if (w[0] != 0)
                    for (uint32_t c = 0; c < 256; ++c) {
                        const uint32_t ci = w[0] ^ c;
                        if (ci < n && (p->opos_ch[ci] & 0xff) == c) sig |= 1u << (c & 31);
                    }
                uint32_t* r = &img->crec[s * 4];
                const uint32_t opos = p->opos_ch[s] >> 8;
                const uint32_t f = w[1];
                uint32_t flags = opos ? 1u : 0u;
                uint32_t fbase = w[2] & 0x7fffffffu;
                if (!lm) {
                    if (w[2] & 0x80000000u) flags |= 2u;  // CF_F2ROOT
                    if (f == kRoot) flags |= 8u, fbase = img->root_base;  // CF_FROOT: fbase pre-resolved to BASE(ROOT)
                } else if (f != kRoot && f != kDead) {
                    const uint32_t f2 = skip_leaves(p->fail[f]);
                    if (f2 == kRoot) flags |= 2u;  // CF_F2ROOT
                    if (f2 == kDead) flags |= 4u;  // CF_F2DEAD
                }
                r[0] = (w[0] << 8) | (p->opos_ch[s] & 0xff);
                r[1] = (f << 8) | flags;
                r[2] = fbase << 8;
                r[3] = sig;
                img->opos_tab[s] = opos;
            }
        }
    } else {
        for (size_t s = 0; s < n; ++s) {
            uint32_t* r = &img->rec[s * 4];
            r[0] = p->base[s];
            r[1] = p->check[s];
            r[2] = p->fail[s];
            r[3] = p->output_pos[s];
        }
        img->root_table.assign(256, kRoot);  // unused by the charwise kernels
        img->mapper = p->mapper_table;
        // compact image for the charwise lane machine (scan_lane.cuh, CwMachine)
        img->root_base = n ? p->base[kRoot] : 0;
        if (n < (size_t(1) << 24) && p->outputs.size() < (size_t(1) << 24)) {
            auto skip_leaves = [&](uint32_t f) {  // failure target with child-less states skipped
                while (f != kRoot && f != kDead && p->base[f] == 0) f = p->fail[f];
                return f;
            };
            // child signatures: bit (mapped code & 15) of the parent for every occupied slot
            std::vector<uint32_t> sig(n, 0);
            for (size_t s = 0; s < n; ++s) {
                const uint32_t par = p->check[s];
                if (par >= n || p->base[par] == 0) continue;
                const uint32_t code = uint32_t(s) ^ p->base[par];
                if (code < p->alphabet_size) sig[par] |= 1u << (code & 15);
            }
            img->crec.resize(n * 4);
            img->opos_tab.resize(n);
            for (size_t s = 0; s < n; ++s) {
                const uint32_t f = skip_leaves(p->fail[s]);
                const bool terminal = f == kRoot || f == kDead;
                uint32_t flags = p->output_pos[s] ? 1u : 0u;
                if (!terminal) {
                    const uint32_t f2 = skip_leaves(p->fail[f]);
                    if (f2 == kRoot) flags |= 2u;  // CF_F2ROOT
                    if (f2 == kDead) flags |= 4u;  // CF_F2DEAD
                }
                const uint32_t chk = p->check[s] < n ? p->check[s] : 0xffffffu;  // vacant: matches no state id
                uint32_t* r = &img->crec[s * 4];
                r[0] = (p->base[s] << 8) | (sig[s] & 0xffu);
                r[1] = (f << 8) | flags;
                r[2] = ((terminal ? 0u : p->base[f]) << 8) | (sig[s] >> 8);
                r[3] = chk << 8;
                img->opos_tab[s] = p->output_pos[s];
            }
        }
    }
    img->outputs.resize(p->outputs.size() * 4);
    for (size_t i = 0; i < p->outputs.size(); ++i) {
        img->outputs[i * 4 + 0] = p->outputs[i].value;
        img->outputs[i * 4 + 1] = p->outputs[i].length;
        img->outputs[i * 4 + 2] = p->outputs[i].parent;
        img->outputs[i * 4 + 3] = 0;
    }
    return DACH_OK;
}

}  // namespace dach
