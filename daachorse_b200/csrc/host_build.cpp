// Host-side automaton construction for libdaachorse_b200.
//
// Construction stays on the host (BASELINE.json north_star).  With a Rust toolchain the
// crate itself builds the automaton and hands its serialize() bytes to
// dach_pma_deserialize(); this file is the host side ABOVE the C ABI for environments
// without Rust.  It produces the same double array, outputs and mapper as the crate
// (so dach_pma_serialize() is byte-identical to the crate's serialize()), which requires
// following the reference's placement rules exactly:
//   trie + failure links + merged outputs ... src/nfa_builder.rs:78-222 (edge order: src/edge_map.rs)
//   vacant-slot ring ........................ src/build_helper.rs:24-227
//   bytewise placement ...................... src/bytewise/builder.rs:204-400
//   charwise placement + code mapper ........ src/charwise/builder.rs:178-359, src/charwise/mapper.rs:16-34
// The engineering differs from the crate: the trie keeps its edges in one open-addressed
// table keyed by (parent, label) and is frozen into a label-sorted CSR before the BFS,
// instead of one sorted edge vector per state.
#include <algorithm>
#include <cstring>
#include <memory>
#include <numeric>

#include "host.h"

namespace dach {

namespace {

// ---- (parent, label) -> child table ------------------------------------------------

class EdgeTable {
  public:
    EdgeTable() { rehash(1u << 12); }

    uint32_t find(uint32_t parent, uint32_t label) const {
        const uint64_t key = pack(parent, label);
        size_t i = slot(key);
        for (;;) {
            if (keys_[i] == key) return vals_[i];
            if (keys_[i] == kEmpty) return kNone;
            i = (i + 1) & mask_;
        }
    }

    void insert(uint32_t parent, uint32_t label, uint32_t child) {
        if ((count_ + 1) * 10 > (mask_ + 1) * 6) rehash((mask_ + 1) * 2);
        put(pack(parent, label), child);
        ++count_;
    }

    static constexpr uint32_t kNone = 0xffffffffu;

  private:
    static constexpr uint64_t kEmpty = ~0ull;
    static uint64_t pack(uint32_t parent, uint32_t label) { return (uint64_t(parent) << 32) | label; }
    size_t slot(uint64_t key) const {
        uint64_t x = key * 0x9e3779b97f4a7c15ull;
        x ^= x >> 29;
        return size_t(x) & mask_;
    }
    void put(uint64_t key, uint32_t v) {
        size_t i = slot(key);
        while (keys_[i] != kEmpty) i = (i + 1) & mask_;
        keys_[i] = key;
        vals_[i] = v;
    }
    void rehash(size_t cap) {
        std::vector<uint64_t> ok;
        std::vector<uint32_t> ov;
        ok.swap(keys_);
        ov.swap(vals_);
        keys_.assign(cap, kEmpty);
        vals_.assign(cap, 0);
        mask_ = cap - 1;
        for (size_t i = 0; i < ok.size(); ++i)
            if (ok[i] != kEmpty) put(ok[i], ov[i]);
    }
    std::vector<uint64_t> keys_;
    std::vector<uint32_t> vals_;
    size_t mask_ = 0, count_ = 0;
};

// ---- sparse trie with failure links and merged output lists -------------------------

struct Edge {
    uint32_t label, child;
};

class SparseTrie {
  public:
    explicit SparseTrie(uint8_t kind) : kind_(kind) {
        // state 0 = root, state 1 = dead (src/nfa_builder.rs:65-75)
        own_head_.assign(2, kNil);
    }

    // NfaBuilder::add (src/nfa_builder.rs:78-113)
    int add(const uint32_t* labels, size_t n, uint64_t byte_len, uint32_t value) {
        if (byte_len > 0xffffffffull) {
            set_error("pattern.len() must be <= u32::MAX");
            return DACH_INVALID_ARGUMENT;
        }
        uint32_t s = kRoot;
        for (size_t i = 0; i < n; ++i) {
            if (kind_ == DACH_LEFTMOST_FIRST && own_head_[s] != kNil) return DACH_OK;  // :87-92
            uint32_t nx = edges_.find(s, labels[i]);
            if (nx == EdgeTable::kNone) {
                if (own_head_.size() > 0xffffffffull) {
                    set_error("state_id must be <= u32::MAX");
                    return DACH_AUTOMATON_SCALE;
                }
                nx = uint32_t(own_head_.size());
                edges_.insert(s, labels[i], nx);
                made_parent_.push_back(s);
                made_label_.push_back(labels[i]);
                own_head_.push_back(kNil);
            }
            s = nx;
        }
        // newest-first chain of this state's own (value, length) pairs
        own_.push_back({value, uint32_t(byte_len), own_head_[s]});
        own_head_[s] = uint32_t(own_.size() - 1);
        ++num_patterns_;
        return DACH_OK;
    }

    size_t num_states() const { return own_head_.size(); }
    size_t num_patterns() const { return num_patterns_; }

    // Freeze the edges into a CSR whose per-state slices are label-ascending -- the
    // iteration order of EdgeMap (src/edge_map.rs:35-43, 69-75) that fixes BFS/DFS order.
    void freeze() {
        const size_t ns = num_states();
        first_.assign(ns + 1, 0);
        for (uint32_t p : made_parent_) ++first_[p + 1];
        for (size_t i = 0; i < ns; ++i) first_[i + 1] += first_[i];
        csr_.resize(made_parent_.size());
        std::vector<uint32_t> cur(first_.begin(), first_.end() - 1);
        for (size_t k = 0; k < made_parent_.size(); ++k)
            csr_[cur[made_parent_[k]]++] = Edge{made_label_[k], uint32_t(k + 2)};
        for (size_t s = 0; s < ns; ++s) {
            Edge* b = csr_.data() + first_[s];
            Edge* e = csr_.data() + first_[s + 1];
            if (e - b > 1) std::sort(b, e, [](const Edge& x, const Edge& y) { return x.label < y.label; });
        }
        made_parent_.clear();
        made_parent_.shrink_to_fit();
        made_label_.clear();
        made_label_.shrink_to_fit();
        fail_.assign(ns, kRoot);
        opos_.assign(ns, 0);
    }

    const Edge* edges_begin(uint32_t s) const { return csr_.data() + first_[s]; }
    const Edge* edges_end(uint32_t s) const { return csr_.data() + first_[s + 1]; }
    bool has_own_output(uint32_t s) const { return own_head_[s] != kNil; }
    uint32_t fail(uint32_t s) const { return fail_[s]; }
    uint32_t output_pos(uint32_t s) const { return opos_[s]; }
    std::vector<OutputRec>& outputs() { return outputs_; }

    // build_fails (src/nfa_builder.rs:115-144) / build_fails_leftmost (:146-201).
    // Returns the breadth-first order.
    std::vector<uint32_t> link_failures() {
        const bool leftmost = is_leftmost(kind_);
        std::vector<uint32_t> order;
        order.reserve(num_states());
        for (const Edge* e = edges_begin(kRoot); e != edges_end(kRoot); ++e) order.push_back(e->child);
        if (leftmost && has_own_output(kRoot))  // :151-160
            for (const Edge* e = edges_begin(kRoot); e != edges_end(kRoot); ++e) fail_[e->child] = kDead;
        for (size_t qi = 0; qi < order.size(); ++qi) {
            const uint32_t s = order[qi];
            if (leftmost && has_own_output(s)) fail_[s] = kDead;  // :169-172
            for (const Edge* e = edges_begin(s); e != edges_end(s); ++e) {
                uint32_t f = fail_[s];
                uint32_t nf;
                if (leftmost && f == kDead) {
                    nf = kDead;  // :177-179
                } else {
                    for (;;) {
                        const uint32_t via = edges_.find(f, e->label);
                        if (via != EdgeTable::kNone) {
                            nf = via;
                            break;
                        }
                        const uint32_t up = fail_[f];
                        if (leftmost && up == kDead) {
                            nf = kDead;
                            break;
                        }
                        if (f == kRoot && up == kRoot) {
                            nf = kRoot;
                            break;
                        }
                        f = up;
                    }
                }
                fail_[e->child] = nf;
                order.push_back(e->child);
            }
        }
        return order;
    }

    // build_outputs (src/nfa_builder.rs:203-222): root first, then BFS order; a state's own
    // patterns are pushed in reverse registration order, each pointing at the previous tail.
    void merge_outputs(const std::vector<uint32_t>& order) {
        outputs_.reserve(own_.size());
        auto emit_state = [&](uint32_t s, uint32_t last) {
            for (uint32_t k = own_head_[s]; k != kNil; k = own_[k].next) {
                outputs_.push_back({own_[k].value, own_[k].length, last});
                last = uint32_t(outputs_.size());
            }
            opos_[s] = last;
        };
        emit_state(kRoot, 0);
        for (uint32_t s : order) emit_state(s, opos_[fail_[s]]);
    }

  private:
    static constexpr uint32_t kNil = 0xffffffffu;
    struct Own {
        uint32_t value, length, next;
    };
    uint8_t kind_;
    EdgeTable edges_;
    std::vector<uint32_t> made_parent_, made_label_;  // edge k created state k + 2
    std::vector<uint32_t> own_head_;
    std::vector<Own> own_;
    size_t num_patterns_ = 0;
    std::vector<uint32_t> first_;
    std::vector<Edge> csr_;
    std::vector<uint32_t> fail_, opos_;
    std::vector<OutputRec> outputs_;
};

// ---- vacancy ring (BuildHelper, src/build_helper.rs) ---------------------------------

class VacancyRing {
  public:
    // BuildHelper::new (:30-44)
    int init(uint32_t block_len, uint32_t num_free_blocks) {
        const uint64_t cap = uint64_t(block_len) * num_free_blocks;
        if (cap > 0xffffffffull) {
            set_error("block_len * num_free_blocks must be <= u32::MAX");
            return DACH_AUTOMATON_SCALE;
        }
        block_len_ = block_len;
        keep_ = num_free_blocks;
        cap_ = uint32_t(cap);
        try {
            next_.assign(cap_, 0);
            prev_.assign(cap_, 0);
            flags_.assign(cap_, 0);
        } catch (const std::bad_alloc&) {
            set_error("out of memory for num_free_blocks");
            return DACH_AUTOMATON_SCALE;
        }
        return DACH_OK;
    }

    uint32_t size() const { return blocks_ * block_len_; }
    uint32_t first_active_block() const { return blocks_ > keep_ ? blocks_ - keep_ : 0; }
    uint32_t block_count() const { return blocks_; }
    bool has_vacant() const { return has_head_; }
    uint32_t head() const { return head_; }
    uint32_t next_of(uint32_t i) const { return next_[at(i)]; }

    bool base_used(uint32_t b) const { return flags_[at(b)] & kBaseBit; }
    bool index_used(uint32_t i) const { return flags_[at(i)] & kIndexBit; }
    void take_base(uint32_t b) { flags_[at(b)] |= kBaseBit; }

    // use_index (:118-130)
    void take_index(uint32_t i) {
        const uint32_t o = at(i);
        flags_[o] |= kIndexBit;
        const uint32_t nx = next_[o], pv = prev_[o];
        next_[at(pv)] = nx;
        prev_[at(nx)] = pv;
        if (head_ == i) {
            if (nx != i)
                head_ = nx;
            else
                has_head_ = false;
        }
    }

    // dropped_block (:177-179)
    bool will_drop(uint32_t* blk) const {
        if (cap_ <= size()) {
            *blk = first_active_block();
            return true;
        }
        return false;
    }

    // push_block (:133-173)
    int grow() {
        if (size() > 0xffffffffu - block_len_) {
            set_error("num_elements must be <= u32::MAX");
            return DACH_AUTOMATON_SCALE;
        }
        uint32_t closing;
        if (will_drop(&closing)) {
            const uint32_t end = (closing + 1) * block_len_;
            while (has_head_ && head_ < end) take_index(head_);
        }
        const uint32_t lo = size(), hi = lo + block_len_;
        ++blocks_;
        for (uint32_t i = lo; i < hi; ++i) {
            const uint32_t o = at(i);
            flags_[o] = 0;
            next_[o] = i + 1;
            prev_[o] = i - 1;
        }
        if (has_head_) {
            const uint32_t tail = prev_[at(head_)];
            prev_[at(lo)] = tail;
            next_[at(tail)] = lo;
            next_[at(hi - 1)] = head_;
            prev_[at(head_)] = hi - 1;
        } else {
            prev_[at(lo)] = hi - 1;
            next_[at(hi - 1)] = lo;
            has_head_ = true;
            head_ = lo;
        }
        return DACH_OK;
    }

    // unused_base_in_block (:76-80): lowest one; returns false if every base is taken
    bool lowest_free_base(uint32_t blk, uint32_t* out) const {
        const uint32_t lo = blk * block_len_;
        for (uint32_t b = lo; b < lo + block_len_; ++b)
            if (!base_used(b)) {
                *out = b;
                return true;
            }
        return false;
    }

  private:
    static constexpr uint8_t kBaseBit = 1, kIndexBit = 2;
    uint32_t at(uint32_t i) const { return i % cap_; }  // callers stay inside the active blocks
    std::vector<uint32_t> next_, prev_;
    std::vector<uint8_t> flags_;
    uint32_t block_len_ = 0, keep_ = 0, blocks_ = 0, cap_ = 0, head_ = 0;
    bool has_head_ = false;
};

// ---- UTF-8 ------------------------------------------------------------------------------

// Strict validation (what Rust's &str guarantees) + decode to code points.
bool decode_utf8(const uint8_t* s, size_t n, std::vector<uint32_t>* out) {
    out->clear();
    size_t i = 0;
    while (i < n) {
        const uint32_t b0 = s[i];
        if (b0 < 0x80) {
            out->push_back(b0);
            ++i;
            continue;
        }
        size_t need;
        uint32_t cp, lo;
        if (b0 >= 0xc2 && b0 <= 0xdf) {
            need = 1, cp = b0 & 0x1f, lo = 0x80;
        } else if (b0 >= 0xe0 && b0 <= 0xef) {
            need = 2, cp = b0 & 0x0f, lo = 0x800;
        } else if (b0 >= 0xf0 && b0 <= 0xf4) {
            need = 3, cp = b0 & 0x07, lo = 0x10000;
        } else {
            return false;
        }
        if (i + need >= n) return false;  // continuation bytes at i+1 .. i+need
        for (size_t k = 1; k <= need; ++k) {
            const uint32_t b = s[i + k];
            if ((b & 0xc0) != 0x80) return false;
            cp = (cp << 6) | (b & 0x3f);
        }
        if (cp < lo || cp > 0x10ffff || (cp >= 0xd800 && cp <= 0xdfff)) return false;
        out->push_back(cp);
        i += need + 1;
    }
    return true;
}

// ---- bytewise placement (src/bytewise/builder.rs:267-400) -----------------------------

class BytewisePlacer {
  public:
    BytewisePlacer(dach_pma* p, const SparseTrie& t) : p_(p), t_(t) {}

    int run(uint32_t num_free_blocks) {
        // init_array (:336-344)
        grow_slots(256);
        int rc = ring_.init(256, num_free_blocks);
        if (rc) return rc;
        rc = ring_.grow();
        if (rc) return rc;
        ring_.take_index(kRoot);
        ring_.take_index(kDead);

        const size_t ns = t_.num_states();
        std::vector<uint32_t> slot_of(ns, kDead);
        slot_of[kRoot] = kRoot;
        std::vector<uint32_t> todo{kRoot};
        uint32_t labels[256];
        while (!todo.empty()) {  // :277-305 (children pushed ascending, so popped largest-first)
            const uint32_t s = todo.back();
            todo.pop_back();
            const Edge* eb = t_.edges_begin(s);
            const Edge* ee = t_.edges_end(s);
            if (eb == ee) continue;
            const size_t ne = size_t(ee - eb);
            for (size_t i = 0; i < ne; ++i) labels[i] = eb[i].label;
            const uint32_t base = pick_base(labels, ne);
            if (size_t(base) >= p_->base.size()) {
                rc = extend();
                if (rc) return rc;
            }
            for (const Edge* e = eb; e != ee; ++e) {
                const uint32_t ci = base ^ e->label;
                ring_.take_index(ci);
                set_check(ci, e->label);
                slot_of[e->child] = ci;
                todo.push_back(e->child);
            }
            p_->base[slot_of[s]] = base;
            ring_.take_base(base);
        }
        // :308-326 fail + output_pos
        for (size_t i = 0; i < ns; ++i) {
            if (i == kDead) continue;
            const uint32_t slot = slot_of[i];
            const uint32_t op = t_.output_pos(uint32_t(i));
            if (op > kU24Max) {  // set_output_pos (src/bytewise.rs:1179-1187)
                set_error("output_pos must be <= 2^24-1");
                return DACH_AUTOMATON_SCALE;
            }
            p_->opos_ch[slot] = (op << 8) | (p_->opos_ch[slot] & 0xff);
            const uint32_t f = t_.fail(uint32_t(i));
            p_->fail[slot] = f == kDead ? kDead : slot_of[f];
        }
        // :328-330
        for (uint32_t b = ring_.first_active_block(); b < ring_.block_count(); ++b) neutralise(b);
        return DACH_OK;
    }

  private:
    void grow_slots(size_t n) {
        p_->base.resize(n, 0);
        p_->fail.resize(n, 0);
        p_->opos_ch.resize(n, 0);
    }
    void set_check(uint32_t i, uint32_t c) { p_->opos_ch[i] = (p_->opos_ch[i] & 0xffffff00u) | (c & 0xff); }

    // find_base / check_valid_base (:347-370)
    uint32_t pick_base(const uint32_t* labels, size_t n) const {
        if (ring_.has_vacant()) {
            uint32_t v = ring_.head();
            do {
                const uint32_t base = v ^ labels[0];
                bool ok = base != 0 && !ring_.base_used(base);
                for (size_t i = 0; ok && i < n; ++i) ok = !ring_.index_used(base ^ labels[i]);
                if (ok) return base;
                v = ring_.next_of(v);
            } while (v != ring_.head());
        }
        return uint32_t(p_->base.size());
    }

    // extend_array (:372-388)
    int extend() {
        if (p_->base.size() > size_t(0xffffffffu - 256u)) {
            set_error("states.len() must be <= u32::MAX");
            return DACH_AUTOMATON_SCALE;
        }
        uint32_t closing;
        if (ring_.will_drop(&closing)) neutralise(closing);
        const int rc = ring_.grow();
        if (rc) return rc;
        grow_slots(p_->base.size() + 256);
        return DACH_OK;
    }

    // remove_invalid_checks (:391-400)
    void neutralise(uint32_t blk) {
        uint32_t ub;
        if (!ring_.lowest_free_base(blk, &ub)) return;
        for (uint32_t c = 0; c < 256; ++c) {
            const uint32_t i = ub ^ c;
            if (i == kRoot || i == kDead || !ring_.index_used(i)) set_check(i, c);
        }
    }

    dach_pma* p_;
    const SparseTrie& t_;
    VacancyRing ring_;
};

// ---- charwise placement (src/charwise/builder.rs:241-359) -----------------------------

uint32_t next_pow2(uint32_t x) {
    uint32_t r = 1;
    while (r < x) r <<= 1;
    return r;
}

class CharwisePlacer {
  public:
    CharwisePlacer(dach_pma* p, const SparseTrie& t) : p_(p), t_(t) {}

    int run(uint32_t num_free_blocks) {
        block_len_ = std::max<uint32_t>(2, next_pow2(p_->alphabet_size));  // init_array (:308-318)
        grow_slots(block_len_);
        int rc = ring_.init(block_len_, num_free_blocks);
        if (rc) return rc;
        rc = ring_.grow();
        if (rc) return rc;
        ring_.take_index(kRoot);
        ring_.take_index(kDead);

        const size_t ns = t_.num_states();
        std::vector<uint32_t> slot_of(ns, kDead);
        slot_of[kRoot] = kRoot;
        std::vector<uint32_t> todo{kRoot};
        std::vector<Edge> coded;
        while (!todo.empty()) {  // :251-281
            const uint32_t s = todo.back();
            todo.pop_back();
            const Edge* eb = t_.edges_begin(s);
            const Edge* ee = t_.edges_end(s);
            if (eb == ee) continue;
            coded.clear();
            for (const Edge* e = eb; e != ee; ++e) coded.push_back({p_->mapper_table[e->label], e->child});
            std::sort(coded.begin(), coded.end(), [](const Edge& a, const Edge& b) { return a.label < b.label; });
            const uint32_t base = pick_base(coded);
            if (p_->base.size() <= size_t(base)) {
                // extend_array (:346-359)
                if (p_->base.size() > size_t(0xffffffffu - block_len_)) {
                    set_error("states.len() must be <= u32::MAX");
                    return DACH_AUTOMATON_SCALE;
                }
                rc = ring_.grow();
                if (rc) return rc;
                grow_slots(p_->base.size() + block_len_);
            }
            const uint32_t parent_slot = slot_of[s];
            for (const Edge& e : coded) {
                const uint32_t ci = base ^ e.label;
                ring_.take_index(ci);
                p_->check[ci] = parent_slot;
                slot_of[e.child] = ci;
                todo.push_back(e.child);
            }
            p_->base[parent_slot] = base;
        }
        for (size_t i = 0; i < ns; ++i) {  // :284-302
            if (i == kDead) continue;
            const uint32_t slot = slot_of[i];
            p_->output_pos[slot] = t_.output_pos(uint32_t(i));
            const uint32_t f = t_.fail(uint32_t(i));
            p_->fail[slot] = f == kDead ? kDead : slot_of[f];
        }
        return DACH_OK;
    }

  private:
    void grow_slots(size_t n) {  // State::default (src/charwise.rs:1103-1112)
        p_->base.resize(n, 0);
        p_->check.resize(n, kDead);
        p_->fail.resize(n, kDead);
        p_->output_pos.resize(n, 0);
    }
    // find_base / verify_base (:320-344)
    uint32_t pick_base(const std::vector<Edge>& coded) const {
        if (ring_.has_vacant()) {
            uint32_t v = ring_.head();
            do {
                const uint32_t base = v ^ coded[0].label;
                bool ok = base != 0;
                for (size_t i = 0; ok && i < coded.size(); ++i) ok = !ring_.index_used(base ^ coded[i].label);
                if (ok) return base;
                v = ring_.next_of(v);
            } while (v != ring_.head());
        }
        return uint32_t(p_->base.size()) ^ coded[0].label;
    }

    dach_pma* p_;
    const SparseTrie& t_;
    VacancyRing ring_;
    uint32_t block_len_ = 2;
};

// CodeMapper::new (src/charwise/mapper.rs:16-34): dense codes by descending frequency,
// ties by ascending code point.
void make_mapper(dach_pma* p, const std::vector<uint32_t>& freqs) {
    std::vector<uint32_t> used;
    for (size_t c = 0; c < freqs.size(); ++c)
        if (freqs[c] != 0) used.push_back(uint32_t(c));
    std::sort(used.begin(), used.end(), [&](uint32_t a, uint32_t b) {
        if (freqs[a] != freqs[b]) return freqs[a] > freqs[b];
        return a < b;
    });
    p->mapper_table.assign(freqs.size(), kInvalidCode);
    for (size_t i = 0; i < used.size(); ++i) p->mapper_table[used[i]] = uint32_t(i);
    p->alphabet_size = uint32_t(used.size());
}

}  // namespace

// build_root_table (src/bytewise.rs:1040-1056)
void rebuild_root_table(dach_pma* p) {
    p->root_table.assign(256, kRoot);
    if (p->base.empty() || p->base[kRoot] == 0) return;
    const uint32_t b = p->base[kRoot];
    for (uint32_t c = 0; c < 256; ++c) {
        const uint32_t ci = b ^ c;
        if (ci < p->base.size() && (p->opos_ch[ci] & 0xff) == c) p->root_table[c] = ci;
    }
}

int build_automaton(bool charwise, const uint8_t* bytes, const uint64_t* offs, const uint32_t* values,
                    uint32_t n, uint8_t match_kind, uint32_t num_free_blocks, dach_pma** out) {
    *out = nullptr;
    if (match_kind > 2) {
        set_error("match_kind must be 0, 1 or 2");
        return DACH_INVALID_ARGUMENT;
    }
    if (num_free_blocks == 0) {  // assert!(n >= 1), src/bytewise/builder.rs:113
        set_error("num_free_blocks must be >= 1");
        return DACH_INVALID_ARGUMENT;
    }
    if (n > 0 && (!offs || (!bytes && offs[n] != 0))) {
        set_error("null pattern buffer");
        return DACH_INVALID_ARGUMENT;
    }
    std::unique_ptr<dach_pma> p(new dach_pma());
    p->charwise = charwise;
    p->match_kind = match_kind;

    SparseTrie trie(match_kind);
    std::vector<uint32_t> labels, freqs;
    for (uint32_t i = 0; i < n; ++i) {
        if (offs[i + 1] < offs[i]) {
            set_error("pattern offsets must be ascending");
            return DACH_INVALID_ARGUMENT;
        }
        const uint8_t* pat = bytes + offs[i];
        const size_t len = size_t(offs[i + 1] - offs[i]);
        if (!charwise) {
            labels.assign(pat, pat + len);
        } else if (!decode_utf8(pat, len, &labels)) {
            set_error("pattern is not valid UTF-8");
            return DACH_INVALID_ARGUMENT;
        }
        const int rc = trie.add(labels.data(), labels.size(), len, values ? values[i] : i);
        if (rc) return rc;
        if (charwise)  // src/charwise/builder.rs:222-228 (counted even for pruned patterns)
            for (uint32_t c : labels) {
                if (freqs.size() <= c) freqs.resize(size_t(c) + 1, 0);
                ++freqs[c];
            }
    }
    if (!charwise && trie.num_patterns() > kU24Max) {  // src/bytewise/builder.rs:256-258
        set_error("patvals.len() must be <= 2^24-1");
        return DACH_AUTOMATON_SCALE;
    }
    if (charwise) make_mapper(p.get(), freqs);
    trie.freeze();
    {
        const std::vector<uint32_t> order = trie.link_failures();
        trie.merge_outputs(order);
    }
    int rc;
    if (charwise) {
        CharwisePlacer placer(p.get(), trie);
        rc = placer.run(num_free_blocks);
    } else {
        BytewisePlacer placer(p.get(), trie);
        rc = placer.run(num_free_blocks);
    }
    if (rc) return rc;
    if (trie.num_states() - 1 > 0xffffffffull) {
        set_error("num_states must be <= u32::MAX");
        return DACH_AUTOMATON_SCALE;
    }
    p->num_states = uint32_t(trie.num_states() - 1);  // -1 for the dead state
    p->outputs.swap(trie.outputs());
    if (!charwise && !is_leftmost(match_kind)) rebuild_root_table(p.get());
    *out = p.release();
    return DACH_OK;
}

}  // namespace dach
