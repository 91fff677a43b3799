// Internal host-side types of libdaachorse_b200 (not part of the C ABI).
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/daachorse_b200.h"

namespace dach {

constexpr uint32_t kRoot = 0;  // ROOT_STATE_IDX (src/bytewise.rs:25, src/charwise.rs)
constexpr uint32_t kDead = 1;  // DEAD_STATE_IDX (src/bytewise.rs:27)
constexpr uint32_t kU24Max = 0x00ffffffu;       // src/intpack.rs:15
constexpr uint32_t kInvalidCode = 0xffffffffu;  // src/charwise/mapper.rs:7

// Output<u32> (src/lib.rs:213-218); parent: 0 = None, else 1-based index.
struct OutputRec {
    uint32_t value, length, parent;
};

void set_error(const std::string& msg);

inline bool is_leftmost(uint8_t k) { return k == DACH_LEFTMOST_LONGEST || k == DACH_LEFTMOST_FIRST; }

}  // namespace dach

// The host automaton.  One struct serves both variants; which arrays are populated
// follows the crate's structs (src/bytewise.rs:54-68, src/charwise.rs:59-65).
struct dach_pma {
    bool charwise = false;
    uint8_t match_kind = 0;
    uint32_t num_states = 0;
    // Double array, one entry per slot.
    //   bytewise: base / fail / opos_ch (State<u32>, src/bytewise.rs:1131-1137).  For a
    //   leftmost automaton the crate keeps (base, opos_ch) in `leftmost_states` and fail in
    //   `fails`; the same three vectors hold them here and serialisation splits them.
    //   charwise: base / check / fail / output_pos (src/charwise.rs:1096-1101).
    std::vector<uint32_t> base, fail;
    std::vector<uint32_t> opos_ch;            // bytewise only
    std::vector<uint32_t> check, output_pos;  // charwise only
    std::vector<uint32_t> root_table;         // bytewise Standard only (256 entries)
    std::vector<uint32_t> mapper_table;       // charwise CodeMapper.table
    uint32_t alphabet_size = 0;               // charwise CodeMapper.alphabet_size
    std::vector<dach::OutputRec> outputs;

    size_t slots() const { return base.size(); }
    uint32_t state_output_pos(size_t i) const { return charwise ? output_pos[i] : (opos_ch[i] >> 8); }
};

namespace dach {

// host_build.cpp
int build_automaton(bool charwise, const uint8_t* bytes, const uint64_t* offs, const uint32_t* values,
                    uint32_t n, uint8_t match_kind, uint32_t num_free_blocks, dach_pma** out);
void rebuild_root_table(dach_pma* p);

// host_wire.cpp
size_t wire_size(const dach_pma* p);
void wire_write(const dach_pma* p, uint8_t* dst);
int wire_read(const uint8_t* src, size_t len, bool charwise, dach_pma** out, size_t* consumed);

}  // namespace dach
