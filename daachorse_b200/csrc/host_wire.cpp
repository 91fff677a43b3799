// The crate's wire format: the hand-off between a Rust caller and this library.
//
//   bytewise  (src/bytewise.rs:801-820):  Vec<State<u32>> | Vec<State<Empty>> | Vec<u32> fails |
//                                         Vec<Output<u32>> | u8 match_kind | u32 num_states
//   charwise  (src/charwise.rs:831-848):  Vec<State> | Vec<u32> table | u32 alphabet_size |
//                                         Vec<Output<u32>> | u8 match_kind | u32 num_states
//   Vec<T> = u32 LE length + items (src/serializer.rs:106-109); all integers little-endian;
//   Option<NonZeroU32> = u32 with 0 for None (src/serializer.rs:75-91).
// Reading applies the crate's own validation (src/bytewise.rs:892-962,
// src/charwise.rs:912-950) -- the device kernels index these arrays unchecked.
#include <cstring>
#include <memory>

#include "host.h"

namespace dach {

namespace {

struct Writer {
    uint8_t* p;
    void u32(uint32_t x) {
        p[0] = uint8_t(x);
        p[1] = uint8_t(x >> 8);
        p[2] = uint8_t(x >> 16);
        p[3] = uint8_t(x >> 24);
        p += 4;
    }
};

struct Reader {
    const uint8_t* p;
    size_t left;
    bool u32(uint32_t* x) {
        if (left < 4) return false;
        *x = uint32_t(p[0]) | uint32_t(p[1]) << 8 | uint32_t(p[2]) << 16 | uint32_t(p[3]) << 24;
        p += 4;
        left -= 4;
        return true;
    }
    // Vec<S> header with the allocation guard of src/serializer.rs:116-118
    bool vec_len(size_t item_bytes, uint32_t* n) {
        if (!u32(n)) return false;
        return uint64_t(*n) * item_bytes <= left;
    }
};

int bad(const char* why) {
    set_error(std::string("invalid automaton: ") + why);
    return DACH_INVALID_AUTOMATON;
}

}  // namespace

size_t wire_size(const dach_pma* p) {
    const size_t n = p->slots(), no = p->outputs.size();
    if (p->charwise) return 4 + n * 16 + 4 + p->mapper_table.size() * 4 + 4 + 4 + no * 12 + 1 + 4;
    if (is_leftmost(p->match_kind)) return 4 + 4 + n * 8 + 4 + n * 4 + 4 + no * 12 + 1 + 4;
    return 4 + n * 12 + 4 + 4 + 4 + no * 12 + 1 + 4;
}

void wire_write(const dach_pma* p, uint8_t* dst) {
    Writer w{dst};
    const size_t n = p->slots();
    if (!p->charwise) {
        const bool lm = is_leftmost(p->match_kind);
        w.u32(lm ? 0 : uint32_t(n));
        if (!lm)
            for (size_t i = 0; i < n; ++i) {
                w.u32(p->base[i]);
                w.u32(p->fail[i]);
                w.u32(p->opos_ch[i]);
            }
        w.u32(lm ? uint32_t(n) : 0);
        if (lm)
            for (size_t i = 0; i < n; ++i) {
                w.u32(p->base[i]);
                w.u32(p->opos_ch[i]);
            }
        w.u32(lm ? uint32_t(n) : 0);
        if (lm)
            for (size_t i = 0; i < n; ++i) w.u32(p->fail[i]);
    } else {
        w.u32(uint32_t(n));
        for (size_t i = 0; i < n; ++i) {
            w.u32(p->base[i]);
            w.u32(p->check[i]);
            w.u32(p->fail[i]);
            w.u32(p->output_pos[i]);
        }
        w.u32(uint32_t(p->mapper_table.size()));
        for (uint32_t x : p->mapper_table) w.u32(x);
        w.u32(p->alphabet_size);
    }
    w.u32(uint32_t(p->outputs.size()));
    for (const OutputRec& o : p->outputs) {
        w.u32(o.value);
        w.u32(o.length);
        w.u32(o.parent);
    }
    *w.p++ = p->match_kind;
    w.u32(p->num_states);
}

int wire_read(const uint8_t* src, size_t len, bool charwise, dach_pma** out, size_t* consumed) {
    *out = nullptr;
    if (!src && len) {
        set_error("null source");
        return DACH_INVALID_ARGUMENT;
    }
    Reader r{src, len};
    std::unique_ptr<dach_pma> p(new dach_pma());
    p->charwise = charwise;
    uint32_t n = 0;

    auto read_outputs_tail = [&]() -> bool {
        if (!r.vec_len(12, &n)) return false;
        p->outputs.resize(n);
        for (uint32_t i = 0; i < n; ++i)
            if (!r.u32(&p->outputs[i].value) || !r.u32(&p->outputs[i].length) || !r.u32(&p->outputs[i].parent))
                return false;
        if (r.left < 1) return false;
        const uint8_t mk = *r.p;
        ++r.p;
        --r.left;
        p->match_kind = (mk == 1 || mk == 2) ? mk : 0;  // From<u8> for MatchKind (src/lib.rs:362-370)
        return r.u32(&p->num_states);
    };

    if (!charwise) {
        uint32_t n_std = 0, n_lm = 0, n_fails = 0;
        std::vector<uint32_t> sb, sf, so, lb, lo, lf;
        if (!r.vec_len(12, &n_std)) return bad("states");
        sb.resize(n_std), sf.resize(n_std), so.resize(n_std);
        for (uint32_t i = 0; i < n_std; ++i)
            if (!r.u32(&sb[i]) || !r.u32(&sf[i]) || !r.u32(&so[i])) return bad("states");
        if (!r.vec_len(8, &n_lm)) return bad("leftmost_states");
        lb.resize(n_lm), lo.resize(n_lm);
        for (uint32_t i = 0; i < n_lm; ++i)
            if (!r.u32(&lb[i]) || !r.u32(&lo[i])) return bad("leftmost_states");
        if (!r.vec_len(4, &n_fails)) return bad("fails");
        lf.resize(n_fails);
        for (uint32_t i = 0; i < n_fails; ++i)
            if (!r.u32(&lf[i])) return bad("fails");
        if (!read_outputs_tail()) return bad("outputs/match_kind/num_states");
        const size_t n_out = p->outputs.size();
        // src/bytewise.rs:892-955
        if (is_leftmost(p->match_kind)) {
            if (n_std != 0) return bad("states must be empty for leftmost");
            if (n_lm == 0) return bad("leftmost_states empty");
            if (n_lm % 256 != 0) return bad("leftmost_states not a multiple of 256");
            if (n_fails != n_lm) return bad("fails length");
            for (uint32_t i = 0; i < n_lm; ++i) {
                if (lb[i] != 0 && lb[i] >= n_lm) return bad("base out of range");
                const uint32_t op = lo[i] >> 8;
                if (op != 0 && size_t(op - 1) >= n_out) return bad("output_pos out of range");
            }
            for (uint32_t i = 0; i < n_lm; ++i)
                if (lf[i] >= n_lm) return bad("fail out of range");
            p->base.swap(lb);
            p->opos_ch.swap(lo);
            p->fail.swap(lf);
        } else {
            if (n_lm != 0 || n_fails != 0) return bad("leftmost arrays must be empty for standard");
            if (n_std == 0) return bad("states empty");
            if (n_std % 256 != 0) return bad("states not a multiple of 256");
            for (uint32_t i = 0; i < n_std; ++i) {
                if (sb[i] != 0 && sb[i] >= n_std) return bad("base out of range");
                if (sf[i] >= n_std) return bad("fail out of range");
                const uint32_t op = so[i] >> 8;
                if (op != 0 && size_t(op - 1) >= n_out) return bad("output_pos out of range");
            }
            p->base.swap(sb);
            p->fail.swap(sf);
            p->opos_ch.swap(so);
            rebuild_root_table(p.get());  // :878-882; its length (256) is the :936 invariant
        }
    } else {
        if (!r.vec_len(16, &n)) return bad("states");
        p->base.resize(n), p->check.resize(n), p->fail.resize(n), p->output_pos.resize(n);
        for (uint32_t i = 0; i < n; ++i)
            if (!r.u32(&p->base[i]) || !r.u32(&p->check[i]) || !r.u32(&p->fail[i]) || !r.u32(&p->output_pos[i]))
                return bad("states");
        if (!r.vec_len(4, &n)) return bad("mapper table");
        p->mapper_table.resize(n);
        for (uint32_t i = 0; i < n; ++i)
            if (!r.u32(&p->mapper_table[i])) return bad("mapper table");
        if (!r.u32(&p->alphabet_size)) return bad("alphabet_size");
        if (!read_outputs_tail()) return bad("outputs/match_kind/num_states");
        // src/charwise.rs:912-943
        for (uint32_t code : p->mapper_table)
            if (code != kInvalidCode && code >= p->alphabet_size) return bad("mapper code out of range");
        size_t block_len = 1;
        while (block_len < p->alphabet_size) block_len <<= 1;
        if (block_len < 2) block_len = 2;
        const size_t ns = p->base.size(), n_out = p->outputs.size();
        if (ns == 0) return bad("states empty");
        if (ns % block_len != 0) return bad("states not a multiple of the block length");
        for (size_t i = 0; i < ns; ++i) {
            if (p->base[i] != 0 && p->base[i] >= ns) return bad("base out of range");
            if (p->fail[i] >= ns) return bad("fail out of range");
            if (p->output_pos[i] != 0 && size_t(p->output_pos[i] - 1) >= n_out) return bad("output_pos out of range");
        }
    }
    // src/bytewise.rs:956-962, src/charwise.rs:944-950: chains strictly decrease
    for (size_t i = 0; i < p->outputs.size(); ++i) {
        const uint32_t parent = p->outputs[i].parent;
        if (parent != 0 && size_t(parent - 1) >= i) return bad("output parent not smaller than its index");
    }
    if (consumed) *consumed = len - r.left;
    *out = p.release();
    return DACH_OK;
}

}  // namespace dach
