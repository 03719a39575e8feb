// Minishogi (5x5 shogi) rules engine: the native counterpart of the external Rust library minishogilib
// v0.6.12 that the reference drives from Python (not vendored, no Rust toolchain).
//
// What it restates, anchored on the reference's call sites and known-answer tests:
//   * Position: set_start_position / set_sfen (incl. "... moves m1 m2 ..."), do_move, generate_moves,
//     get_side_to_move, get_ply                      selfplay.py:11-36,78; usi.py:44-140
//   * is_repetition() -> (is_repetition, my_check_repetition, op_check_repetition)
//                                                     selfplay.py:17-26; tests/test_repetition.py (8 KATs)
//   * solve_checkmate_dfs(depth) -> (found, move)     selfplay.py:36; tests/test_checkmate.py (9 KATs)
//   * move strings in USI/SFEN form ("5e4d", "3d4e+", "G*1d").
// Rules: standard minishogi -- K, G, S, B, R, P per side; promotion on the far rank only; drops with nifu,
// last-rank and pawn-drop-mate restrictions; fourth occurrence of a position = repetition (sennichite),
// perpetual check loses.  The network input encoding and the policy index (to_alphazero_input, and the move
// indexing inside minishogilib.MCTS.evaluate) are NOT visible in the reference tree; they are restated here
// as documented below and are "parity unpinned" against minishogilib.
//
// Representation: a 25-bit bitboard per (colour, piece type) beside the mailbox board.  Legal moves are generated
// directly (checkers, pins and king-danger squares from attack tables) instead of by make/unmake trials; the hash
// is an incrementally updated Zobrist key.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#if defined(__CUDACC__)
#define MS_HD __host__ __device__
#else
#define MS_HD
#endif

namespace mshogi {

constexpr int kN = 5, kSquares = 25;
enum Color : int { BLACK = 0, WHITE = 1 };
enum PieceType : int { EMPTY = 0, KING, GOLD, SILVER, BISHOP, ROOK, PAWN, PRO_SILVER, HORSE, DRAGON, TOKIN, PIECE_TYPES };
constexpr int kHandTypes = 5;  // GOLD, SILVER, BISHOP, ROOK, PAWN
constexpr PieceType kHandPiece[kHandTypes] = {GOLD, SILVER, BISHOP, ROOK, PAWN};
using Bitboard = uint32_t;     // bit sq = rank * 5 + col
constexpr Bitboard kAllSquares = (1u << kSquares) - 1u;

MS_HD inline PieceType hand_piece(int k) {   // hand slot -> type (kHandPiece, callable from device code)
  switch (k) { case 0: return GOLD; case 1: return SILVER; case 2: return BISHOP; case 3: return ROOK; default: return PAWN; }
}
MS_HD inline int hand_index(PieceType t) {  // unpromoted type -> hand slot
  switch (t) { case GOLD: return 0; case SILVER: return 1; case BISHOP: return 2; case ROOK: return 3; case PAWN: return 4; default: return -1; }
}
MS_HD inline PieceType promote(PieceType t) {
  switch (t) { case SILVER: return PRO_SILVER; case BISHOP: return HORSE; case ROOK: return DRAGON; case PAWN: return TOKIN; default: return t; }
}
MS_HD inline PieceType unpromote(PieceType t) {
  switch (t) { case PRO_SILVER: return SILVER; case HORSE: return BISHOP; case DRAGON: return ROOK; case TOKIN: return PAWN; default: return t; }
}
MS_HD inline bool can_promote(PieceType t) { return t == SILVER || t == BISHOP || t == ROOK || t == PAWN; }

// square = rank * 5 + col; rank 0 = 'a' (White's home row), col 0 = file 5 (left edge as SFEN prints it)
MS_HD inline int sq_of(int rank, int col) { return rank * kN + col; }
MS_HD inline int rank_of(int sq) { return sq / kN; }
MS_HD inline int col_of(int sq) { return sq % kN; }
MS_HD inline int lsb(Bitboard b) {
#ifdef __CUDA_ARCH__
  return __ffs(static_cast<int>(b)) - 1;
#else
  return __builtin_ctz(b);
#endif
}
MS_HD inline int msb(Bitboard b) {
#ifdef __CUDA_ARCH__
  return 31 - __clz(static_cast<int>(b));
#else
  return 31 - __builtin_clz(b);
#endif
}
MS_HD inline int pop_lsb(Bitboard& b) { const int s = lsb(b); b &= b - 1; return s; }

struct Move {
  int8_t from = -1;      // -1: drop
  int8_t to = 0;
  int8_t piece = 0;      // PieceType moved (before promotion) or dropped
  int8_t promotes = 0;
  int8_t captured = 0;   // PieceType on `to` (as it stood, promoted or not), filled by generate/do
  MS_HD bool is_drop() const { return from < 0; }
  MS_HD bool operator==(const Move& o) const { return from == o.from && to == o.to && piece == o.piece && promotes == o.promotes; }
  std::string sfen() const {
    static const char* kLetters = " KGSBRP";
    std::string s;
    auto sqs = [](int sq) { std::string t; t += static_cast<char>('5' - col_of(sq)); t += static_cast<char>('a' + rank_of(sq)); return t; };
    if (is_drop()) { s += kLetters[piece]; s += '*'; s += sqs(to); }
    else { s += sqs(from); s += sqs(to); if (promotes) s += '+'; }
    return s;
  }
};

// step directions as (drank, dcol) for BLACK (who moves towards rank 0); WHITE mirrors the rank sign.
// order: N, NE, E, SE, S, SW, W, NW
constexpr int kDirR[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
constexpr int kDirC[8] = {0, 1, 1, 1, 0, -1, -1, -1};

MS_HD inline uint32_t step_mask(PieceType t) {  // bit d set: may step one square in direction d (BLACK orientation)
  switch (t) {
    case KING: return 0xFF;
    case GOLD: case PRO_SILVER: case TOKIN: return (1u << 0) | (1u << 1) | (1u << 7) | (1u << 2) | (1u << 6) | (1u << 4);
    case SILVER: return (1u << 0) | (1u << 1) | (1u << 7) | (1u << 3) | (1u << 5);
    case PAWN: return 1u << 0;
    case HORSE: return (1u << 0) | (1u << 2) | (1u << 4) | (1u << 6);
    case DRAGON: return (1u << 1) | (1u << 3) | (1u << 5) | (1u << 7);
    default: return 0;
  }
}
MS_HD inline uint32_t slide_mask(PieceType t) {
  switch (t) {
    case BISHOP: case HORSE: return (1u << 1) | (1u << 3) | (1u << 5) | (1u << 7);
    case ROOK: case DRAGON: return (1u << 0) | (1u << 2) | (1u << 4) | (1u << 6);
    default: return 0;
  }
}
constexpr uint32_t kRookDirs = (1u << 0) | (1u << 2) | (1u << 4) | (1u << 6);
constexpr uint32_t kBishopDirs = (1u << 1) | (1u << 3) | (1u << 5) | (1u << 7);

// ---- attack tables (built once) ------------------------------------------------------------------------------
struct Tables {
  Bitboard step_to[2][PIECE_TYPES][kSquares];   // squares a (colour, type) piece standing on sq steps to
  Bitboard step_from[2][PIECE_TYPES][kSquares]; // squares from which a (colour, type) piece steps onto sq
  Bitboard ray[8][kSquares];                    // all squares from sq in absolute direction d (N = towards rank 0)
  Bitboard between[kSquares][kSquares];         // squares strictly between two aligned squares (else 0)
  Bitboard line[kSquares][kSquares];            // the whole line through two aligned squares incl. both (else 0)
  Bitboard file_mask[kN];
  Bitboard rank_mask[kN];
  uint64_t zob_piece[kSquares][32];
  uint64_t zob_hand[2][kHandTypes][4];
  uint64_t zob_side[2];

  static uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
  }
  Tables() {
    memset(static_cast<void*>(this), 0, sizeof(*this));
    auto on = [](int r, int c) { return r >= 0 && r < kN && c >= 0 && c < kN; };
    for (int s = 0; s < kSquares; ++s) {
      const int r = rank_of(s), c = col_of(s);
      for (int d = 0; d < 8; ++d) {
        int rr = r + kDirR[d], cc = c + kDirC[d];
        Bitboard acc = 0;
        while (on(rr, cc)) {
          const int t = sq_of(rr, cc);
          between[s][t] = acc;
          acc |= 1u << t;
          rr += kDirR[d]; cc += kDirC[d];
        }
        ray[d][s] = acc;
      }
      for (int color = 0; color < 2; ++color)
        for (int t = 1; t < PIECE_TYPES; ++t) {
          const uint32_t steps = step_mask(static_cast<PieceType>(t));
          for (int d = 0; d < 8; ++d) {
            if (!(steps >> d & 1)) continue;
            const int rr = r + kDirR[d] * (color == BLACK ? 1 : -1), cc = c + kDirC[d];
            if (on(rr, cc)) step_to[color][t][s] |= 1u << sq_of(rr, cc);
          }
        }
      file_mask[c] |= 1u << s;
      rank_mask[r] |= 1u << s;
    }
    for (int color = 0; color < 2; ++color)
      for (int t = 1; t < PIECE_TYPES; ++t)
        for (int s = 0; s < kSquares; ++s)
          for (Bitboard b = step_to[color][t][s]; b;) step_from[color][t][pop_lsb(b)] |= 1u << s;
    for (int a = 0; a < kSquares; ++a)
      for (int d = 0; d < 8; ++d)
        for (Bitboard b = ray[d][a]; b;) line[a][pop_lsb(b)] = ray[d][a] | ray[(d + 4) & 7][a] | (1u << a);
    for (int s = 0; s < kSquares; ++s)
      for (int p = 1; p < 32; ++p) zob_piece[s][p] = mix(static_cast<uint64_t>(s) * 64 + static_cast<uint64_t>(p));
    for (int c = 0; c < 2; ++c)
      for (int k = 0; k < kHandTypes; ++k)
        for (int n = 1; n < 4; ++n) zob_hand[c][k][n] = mix(4096 + static_cast<uint64_t>((c * kHandTypes + k) * 8 + n));
    zob_side[BLACK] = 0x1234567ull; zob_side[WHITE] = 0x7654321ull;
  }
};
#if defined(__CUDACC__)
static __device__ const Tables* g_device_tables = nullptr;   // set once by the library (a copy of the host tables)
#endif
MS_HD inline const Tables& tables() {
#ifdef __CUDA_ARCH__
  return *g_device_tables;
#else
  static const Tables t;
  return t;
#endif
}

// Squares attacked by a slider on `sq` along the directions in `dirs` (bit d), given the occupancy.
MS_HD inline Bitboard slide_attacks(const Tables& T, int sq, Bitboard occ, uint32_t dirs) {
  Bitboard att = 0;
  for (int d = (dirs & 1) ? 0 : 1; d < 8; d += 2) {     // the four rook or the four bishop directions
    const Bitboard r = T.ray[d][sq];
    const Bitboard blockers = r & occ;
    if (!blockers) { att |= r; continue; }
    // directions 2 (E), 3 (SE), 4 (S), 5 (SW) run towards higher square numbers
    const int b = (d >= 2 && d <= 5) ? lsb(blockers) : msb(blockers);
    att |= r ^ T.ray[d][b];
  }
  return att;
}

struct MoveList {                    // no position has anywhere near 256 legal moves on a 5x5 board
  Move m[256];
  int n = 0;
  bool in_check = false;             // set by generate_moves: the side to move is in check
  MS_HD void push(const Move& mv) { if (n < 256) m[n++] = mv; }
  MS_HD const Move* begin() const { return m; }
  MS_HD const Move* end() const { return m + n; }
  MS_HD bool empty() const { return n == 0; }
  MS_HD int size() const { return n; }
};

struct Snapshot {                    // what the input encoder needs of a past position
  Bitboard pieces[2][PIECE_TYPES - 1];   // [colour][type - 1]
  int8_t hand[2][kHandTypes];
  uint8_t rep_count;                // occurrences of that position up to then (1..)
};

struct HistoryEntry {                // one per position of the game so far (index 0 = the position set_sfen gave)
  uint64_t hash;
  Snapshot snap;
  uint8_t gave_check;               // the move leading here gave check (index 0: unused)
  Move move;                        // the move leading here (index 0: unused)
};

// The rules core: board, hands, side to move and everything that needs no game history.  Plain data and plain
// functions, usable from host code and (under nvcc) from device code alike.
struct Board {
  int8_t board[kSquares];            // 0 = empty, else type | (color << 4)
  int8_t hand[2][kHandTypes];
  Bitboard bb[2][PIECE_TYPES];       // [colour][type]; index 0 = all pieces of that colour
  int side = BLACK;
  int ply = 0;
  int8_t king_sq[2] = {-1, -1};      // cached king squares (-1: no king, as in some test positions)

  MS_HD static int type_of(int8_t p) { return p & 15; }
  MS_HD static int color_of(int8_t p) { return (p >> 4) & 1; }
  MS_HD static int8_t make_piece(int type, int color) { return static_cast<int8_t>(type | (color << 4)); }

  MS_HD void clear_board() {
    memset(board, 0, sizeof(board));
    memset(hand, 0, sizeof(hand));
    memset(bb, 0, sizeof(bb));
    side = BLACK; ply = 0;
    king_sq[0] = king_sq[1] = -1;
  }

  // ---- attacks -------------------------------------------------------------------------------
  MS_HD Bitboard occupied() const { return bb[BLACK][0] | bb[WHITE][0]; }

  // Pieces of colour `by` that attack `sq`, with `occ` as the blocking set (pass a modified occupancy to look
  // through a piece that is about to move).
  MS_HD Bitboard attackers_to(int sq, int by, Bitboard occ) const {
    const Tables& T = tables();
    const Bitboard* b = bb[by];
    Bitboard att = (T.step_from[by][PAWN][sq] & b[PAWN]) | (T.step_from[by][SILVER][sq] & b[SILVER]) |
                   (T.step_from[by][GOLD][sq] & (b[GOLD] | b[PRO_SILVER] | b[TOKIN])) |
                   (T.step_from[by][KING][sq] & (b[KING] | b[HORSE] | b[DRAGON]));
    const Bitboard rooks = b[ROOK] | b[DRAGON], bishops = b[BISHOP] | b[HORSE];
    if (rooks) att |= slide_attacks(T, sq, occ, kRookDirs) & rooks;
    if (bishops) att |= slide_attacks(T, sq, occ, kBishopDirs) & bishops;
    return att;
  }

  MS_HD bool attacked_by(int sq, int by) const { return attackers_to(sq, by, occupied()) != 0; }

  MS_HD int king_square(int color) const { return king_sq[color]; }

  MS_HD bool in_check(int color) const {
    const int k = king_square(color);
    return k >= 0 && attacked_by(k, color ^ 1);
  }

  // ---- make / unmake (board + hands + side only; history handled by do_move / undo_move) -------------
  MS_HD void apply(const Move& m) {
    const int us = side;
    if (m.is_drop()) {
      put(m.to, m.piece, us);
      --hand[us][hand_index(static_cast<PieceType>(m.piece))];
    } else {
      if (m.captured) {
        lift(m.to, m.captured, us ^ 1);
        ++hand[us][hand_index(unpromote(static_cast<PieceType>(m.captured)))];
      }
      lift(m.from, m.piece, us);
      put(m.to, m.promotes ? promote(static_cast<PieceType>(m.piece)) : m.piece, us);
    }
    side ^= 1;
  }

  MS_HD void revert(const Move& m) {
    side ^= 1;
    const int us = side;
    if (m.is_drop()) {
      lift(m.to, m.piece, us);
      ++hand[us][hand_index(static_cast<PieceType>(m.piece))];
    } else {
      lift(m.to, m.promotes ? promote(static_cast<PieceType>(m.piece)) : m.piece, us);
      put(m.from, m.piece, us);
      if (m.captured) {
        put(m.to, m.captured, us ^ 1);
        --hand[us][hand_index(unpromote(static_cast<PieceType>(m.captured)))];
      }
    }
  }

  // ---- move generation ------------------------------------------------------------------------
  // Fully legal moves of the side to move (own king never left in check; nifu, last-rank pawn and
  // pawn-drop-mate excluded).
  MS_HD void generate_moves(MoveList& out) { out.n = 0; out.in_check = false; generate<false>(&out); }

  MS_HD bool has_legal_move() { return generate<true>(nullptr); }

  // ---- mate search (tests/test_checkmate.py) ----------------------------------------------------
  // Depth-limited AND/OR search over checking sequences: the side to move must check on every move.
  // node_budget > 0 bounds the work: when that many positions have been looked at the search gives up and reports
  // "no mate" (a few positions with many checking drops have trees of millions of nodes).
  // exhausted (optional) reports that the budget ran out before the search was complete (a "no mate" that is not proven).
  MS_HD bool solve_checkmate_dfs(int depth, Move& best, long node_budget = 0, bool* exhausted = nullptr) {
    long left = node_budget > 0 ? node_budget : -1;
    const bool found = depth > 0 && mate_attack(depth, &best, left);
    if (exhausted) *exhausted = !found && left == 0;
    return found;
  }

  MS_HD uint64_t compute_hash() const {
    const Tables& T = tables();
    uint64_t h = T.zob_side[side];
    for (int s = 0; s < kSquares; ++s) h ^= T.zob_piece[s][board[s] & 31];
    for (int c = 0; c < 2; ++c)
      for (int k = 0; k < kHandTypes; ++k) h ^= T.zob_hand[c][k][hand[c][k] & 3];
    return h;
  }

  MS_HD void put(int sq, int type, int color) {
    board[sq] = make_piece(type, color);
    bb[color][type] |= 1u << sq;
    bb[color][0] |= 1u << sq;
    if (type == KING) king_sq[color] = static_cast<int8_t>(sq);
  }

  MS_HD void lift(int sq, int type, int color) {
    board[sq] = 0;
    bb[color][type] &= ~(1u << sq);
    bb[color][0] &= ~(1u << sq);
  }

  // Key of the position after `m`, given the key of this one (called before apply).
  MS_HD uint64_t hash_after(uint64_t key, const Move& m) const {
    const Tables& T = tables();
    const int us = side;
    uint64_t h = key ^ T.zob_side[BLACK] ^ T.zob_side[WHITE];
    if (m.is_drop()) {
      const int k = hand_index(static_cast<PieceType>(m.piece));
      h ^= T.zob_hand[us][k][hand[us][k] & 3] ^ T.zob_hand[us][k][(hand[us][k] - 1) & 3];
      h ^= T.zob_piece[m.to][make_piece(m.piece, us) & 31];
    } else {
      if (m.captured) {
        const int k = hand_index(unpromote(static_cast<PieceType>(m.captured)));
        h ^= T.zob_piece[m.to][make_piece(m.captured, us ^ 1) & 31];
        h ^= T.zob_hand[us][k][hand[us][k] & 3] ^ T.zob_hand[us][k][(hand[us][k] + 1) & 3];
      }
      h ^= T.zob_piece[m.from][make_piece(m.piece, us) & 31];
      h ^= T.zob_piece[m.to][make_piece(m.promotes ? promote(static_cast<PieceType>(m.piece)) : m.piece, us) & 31];
    }
    return h;
  }

  MS_HD static void emit(MoveList* out, int from, int to, PieceType t, int captured, bool promo_zone, int far_rank) {
    Move m;
    m.from = static_cast<int8_t>(from); m.to = static_cast<int8_t>(to); m.piece = static_cast<int8_t>(t);
    m.captured = static_cast<int8_t>(captured);
    if (promo_zone && can_promote(t)) { Move pm = m; pm.promotes = 1; out->push(pm); }
    if (!(t == PAWN && rank_of(to) == far_rank)) out->push(m);       // a pawn on the last rank must promote
  }

  // Legal move generation.  kAny: stop at the first legal move and return true (`out` is not used).
  // kChecks: only moves that can give check -- a superset of the checking moves, in the order of the full list (the
  // piece lands where its kind attacks the enemy king with our own men thought away, or leaves a line through that
  // king); the mate search confirms each by playing it.
  template <bool kAny, bool kChecks = false>
  MS_HD bool generate(MoveList* out) {
    const Tables& T = tables();
    const int us = side, them = us ^ 1;
    const int far = us == BLACK ? 0 : kN - 1;      // promotion rank
    const Bitboard own = bb[us][0], occ = occupied();
    const int ksq = king_sq[us];
    Bitboard zone[PIECE_TYPES] = {};               // kChecks: squares from which a piece of that type would attack their king
    Bitboard king_lines = 0;                       // kChecks: lines through their king that one of our sliders could use
    if (kChecks) {
      const int ek = king_sq[them];
      if (ek < 0) return false;
      const Bitboard rook_rays = slide_attacks(T, ek, bb[them][0], kRookDirs), bishop_rays = slide_attacks(T, ek, bb[them][0], kBishopDirs);
      zone[GOLD] = zone[PRO_SILVER] = zone[TOKIN] = T.step_from[us][GOLD][ek];
      zone[SILVER] = T.step_from[us][SILVER][ek];
      zone[PAWN] = T.step_from[us][PAWN][ek];
      zone[BISHOP] = bishop_rays; zone[ROOK] = rook_rays;
      zone[HORSE] = bishop_rays | T.step_from[us][KING][ek];
      zone[DRAGON] = rook_rays | T.step_from[us][KING][ek];
      if (bb[us][ROOK] | bb[us][DRAGON]) king_lines |= rook_rays;
      if (bb[us][BISHOP] | bb[us][HORSE]) king_lines |= bishop_rays;
    }
    Bitboard checkers = 0, pinned = 0;
    Bitboard target_mask = kAllSquares;            // where non-king pieces may land (blocks / captures when in check)
    if (ksq >= 0) {
      checkers = attackers_to(ksq, them, occ);
      // pins: an enemy slider aligned with the king with exactly one piece, ours, in between
      const Bitboard rooks = bb[them][ROOK] | bb[them][DRAGON], bishops = bb[them][BISHOP] | bb[them][HORSE];
      Bitboard snipers = 0;
      if (rooks) snipers |= slide_attacks(T, ksq, 0, kRookDirs) & rooks;
      if (bishops) snipers |= slide_attacks(T, ksq, 0, kBishopDirs) & bishops;
      while (snipers) {
        const Bitboard b = T.between[ksq][pop_lsb(snipers)] & occ;
        if (b && !(b & (b - 1)) && (b & own)) pinned |= b;
      }
      if (!kAny) out->in_check = checkers != 0;
      if (checkers) {
        if (checkers & (checkers - 1)) target_mask = 0;                                   // double check: king moves only
        else target_mask = checkers | T.between[ksq][lsb(checkers)];
      }
      // king moves: the destination must not be attacked once the king has left its square
      for (Bitboard dst = (kChecks && !(king_lines >> ksq & 1)) ? 0 : T.step_to[us][KING][ksq] & ~own; dst;) {
        const int to = pop_lsb(dst);
        if (attackers_to(to, them, occ ^ (1u << ksq))) continue;
        if (kAny) return true;
        emit(out, ksq, to, KING, type_of(board[to]), false, far);
      }
    }
    if (target_mask) {
      for (Bitboard pcs = own & ~bb[us][KING]; pcs;) {
        const int from = pop_lsb(pcs);
        const PieceType t = static_cast<PieceType>(type_of(board[from]));
        Bitboard dst = T.step_to[us][t][from];
        const uint32_t slides = slide_mask(t);
        if (slides) dst |= slide_attacks(T, from, occ, slides);
        dst &= ~own & target_mask;
        if (pinned >> from & 1) dst &= T.line[ksq][from];
        if (kChecks && !(king_lines >> from & 1)) dst &= zone[t] | (can_promote(t) ? zone[promote(t)] : 0);
        if (kAny) { if (dst) return true; continue; }    // (a pawn reaching the last rank still has its promotion)
        const bool from_zone = rank_of(from) == far;
        while (dst) {
          const int to = pop_lsb(dst);
          emit(out, from, to, t, type_of(board[to]), from_zone || rank_of(to) == far, far);
        }
      }
      // drops: onto empty squares (only interpositions when in check)
      const Bitboard empties = ~occ & kAllSquares & target_mask;
      if (empties)
        for (int k = 0; k < kHandTypes; ++k) {
          if (!hand[us][k]) continue;
          const PieceType t = hand_piece(k);
          Bitboard dst = kChecks ? empties & zone[t] : empties;
          if (t == PAWN) {
            dst &= ~T.rank_mask[far];                                                       // not onto the last rank
            for (Bitboard p = bb[us][PAWN]; p;) dst &= ~T.file_mask[col_of(pop_lsb(p))];    // nifu
          }
          while (dst) {
            const int to = pop_lsb(dst);
            Move m;
            m.from = -1; m.to = static_cast<int8_t>(to); m.piece = static_cast<int8_t>(t);
            if (t == PAWN && king_sq[them] >= 0 && (T.step_to[us][PAWN][to] >> king_sq[them] & 1)) {
              apply(m);                                                    // pawn-drop check: illegal if it is mate
              const bool mate = !has_legal_move();
              revert(m);
              if (mate) continue;
            }
            if (kAny) return true;
            out->push(m);
          }
        }
    }
    return kAny ? false : out->n > 0;
  }

  MS_HD bool mate_attack(int depth, Move* best, long& left) {
    if (left == 0) return false;
    if (left > 0) --left;
    MoveList moves;
    moves.n = 0;
    generate<false, true>(&moves);
    for (const Move& m : moves) {
      apply(m);
      bool mated = false;
      if (in_check(side)) mated = mate_defend(depth - 1, left);
      revert(m);
      if (mated) { if (best) *best = m; return true; }
    }
    return false;
  }

  // true iff the side to move (in check) is mated within `depth` further plies whatever it plays
  MS_HD bool mate_defend(int depth, long& left) {
    if (left == 0) return false;
    if (left > 0) --left;
    if (depth <= 1) return !has_legal_move();
    MoveList moves;
    generate_moves(moves);
    if (moves.empty()) return true;
    for (const Move& m : moves) {
      apply(m);
      const bool still = mate_attack(depth - 1, nullptr, left);
      revert(m);
      if (!still) return false;
    }
    return true;
  }

};

class Position : public Board {
 public:
  // History for repetition / encoding / tree reuse; entry(ply) = the current position.  A position made by fork_from()
  // shares the other position's entries as a read-only prefix and owns only what it plays on top of them.
  std::vector<HistoryEntry> hist;
  const HistoryEntry* base_ = nullptr;
  int base_n_ = 0;

  Position() { set_start_position(); }
  Position(const Position& o) { *this = o; }
  Position& operator=(const Position& o) {     // a copy owns its whole history
    if (this == &o) return *this;
    copy_board_from(o);
    hist.clear();
    hist.reserve(static_cast<size_t>(o.history_size()) + 8);
    for (int i = 0; i < o.history_size(); ++i) hist.push_back(o.entry(i));
    base_ = nullptr; base_n_ = 0;
    return *this;
  }
  // Cheap copy for a search: valid while `root` is alive and unchanged; `root` must own its history.
  void fork_from(const Position& root) {
    copy_board_from(root);
    hist.clear();
    base_ = root.hist.data();
    base_n_ = static_cast<int>(root.hist.size());
  }
  int history_size() const { return base_n_ + static_cast<int>(hist.size()); }
  const HistoryEntry& entry(int i) const { return i < base_n_ ? base_[i] : hist[static_cast<size_t>(i - base_n_)]; }
  const HistoryEntry& last_entry() const { return hist.empty() ? base_[base_n_ - 1] : hist.back(); }


  void clear() {
    clear_board();
    hist.clear();
    base_ = nullptr; base_n_ = 0;
  }

  void set_start_position() { set_sfen("rbsgk/4p/5/P4/KGSBR b - 1"); }

  // "board side hands movenumber [moves m1 m2 ...]"
  bool set_sfen(const std::string& sfen) {
    clear();
    size_t i = 0;
    int r = 0, c = 0;
    bool promo = false;
    for (; i < sfen.size() && sfen[i] != ' '; ++i) {
      const char ch = sfen[i];
      if (ch == '/') { ++r; c = 0; }
      else if (ch == '+') promo = true;
      else if (ch >= '1' && ch <= '5') c += ch - '0';
      else {
        const int color = (ch >= 'a' && ch <= 'z') ? WHITE : BLACK;
        PieceType t = letter_type(ch);
        if (t == EMPTY || r >= kN || c >= kN) return false;
        if (promo) t = promote(t);
        promo = false;
        put(sq_of(r, c), t, color);
        ++c;
      }
    }
    while (i < sfen.size() && sfen[i] == ' ') ++i;
    if (i >= sfen.size()) return false;
    side = sfen[i] == 'w' ? WHITE : BLACK;
    ++i;
    while (i < sfen.size() && sfen[i] == ' ') ++i;
    int count = 0;
    for (; i < sfen.size() && sfen[i] != ' '; ++i) {
      const char ch = sfen[i];
      if (ch == '-') continue;
      if (ch >= '0' && ch <= '9') { count = count * 10 + (ch - '0'); continue; }
      const int color = (ch >= 'a' && ch <= 'z') ? WHITE : BLACK;
      const int hi = hand_index(letter_type(ch));
      if (hi < 0) return false;
      hand[color][hi] = static_cast<int8_t>(hand[color][hi] + (count ? count : 1));
      count = 0;
    }
    while (i < sfen.size() && sfen[i] == ' ') ++i;
    while (i < sfen.size() && sfen[i] != ' ') ++i;  // move number (the reference always restarts ply at 0)
    push_history(compute_hash(), false, Move());
    // optional move list
    size_t m = sfen.find("moves", i);
    if (m != std::string::npos) {
      m += 5;
      while (m < sfen.size()) {
        while (m < sfen.size() && sfen[m] == ' ') ++m;
        size_t e = m;
        while (e < sfen.size() && sfen[e] != ' ') ++e;
        if (e > m) {
          Move mv;
          if (!parse_move(sfen.substr(m, e - m), mv)) return false;
          do_move(mv);
        }
        m = e;
      }
    }
    return true;
  }

  std::string sfen() const {
    static const char* kLetters = " KGSBRPSBRP";
    std::string s;
    for (int r = 0; r < kN; ++r) {
      int empty = 0;
      for (int c = 0; c < kN; ++c) {
        const int8_t p = board[sq_of(r, c)];
        if (!p) { ++empty; continue; }
        if (empty) { s += static_cast<char>('0' + empty); empty = 0; }
        const int t = type_of(p);
        if (t >= PRO_SILVER) s += '+';
        char ch = kLetters[t];
        if (color_of(p) == WHITE) ch = static_cast<char>(ch - 'A' + 'a');
        s += ch;
      }
      if (empty) s += static_cast<char>('0' + empty);
      if (r + 1 < kN) s += '/';
    }
    s += side == BLACK ? " b " : " w ";
    std::string h;
    static const char kHandLetters[kHandTypes] = {'G', 'S', 'B', 'R', 'P'};
    for (int col = 0; col < 2; ++col)
      for (int k = 0; k < kHandTypes; ++k)
        if (hand[col][k]) {
          if (hand[col][k] > 1) h += static_cast<char>('0' + hand[col][k]);
          h += col == BLACK ? kHandLetters[k] : static_cast<char>(kHandLetters[k] - 'A' + 'a');
        }
    s += h.empty() ? "-" : h;
    s += " " + std::to_string(ply + 1);
    return s;
  }

  std::string pretty() const {       // Position.print(): the board as text, White's home rank on top
    static const char* kLetters = " KGSBRPSBRP";
    std::string s = "  5  4  3  2  1\n";
    for (int r = 0; r < kN; ++r) {
      for (int c = 0; c < kN; ++c) {
        const int8_t p = board[sq_of(r, c)];
        if (!p) { s += "  ."; continue; }
        const int t = type_of(p);
        s += ' ';
        s += t >= PRO_SILVER ? '+' : ' ';
        s += color_of(p) == WHITE ? static_cast<char>(kLetters[t] - 'A' + 'a') : kLetters[t];
      }
      s += "  ";
      s += static_cast<char>('a' + r);
      s += '\n';
    }
    s += "sfen " + sfen() + "\nply " + std::to_string(ply) + (side == BLACK ? " (first player to move)\n" : " (second player to move)\n");
    return s;
  }

  bool parse_move(const std::string& s, Move& mv) const {
    auto sq = [](char f, char r, int& out) {
      if (f < '1' || f > '5' || r < 'a' || r > 'e') return false;
      out = sq_of(r - 'a', '5' - f);
      return true;
    };
    mv = Move();
    if (s.size() >= 4 && s[1] == '*') {
      const PieceType t = letter_type(s[0]);
      int to;
      if (hand_index(t) < 0 || !sq(s[2], s[3], to)) return false;
      mv.from = -1; mv.to = static_cast<int8_t>(to); mv.piece = static_cast<int8_t>(t);
      return true;
    }
    int from, to;
    if (s.size() < 4 || !sq(s[0], s[1], from) || !sq(s[2], s[3], to) || !board[from]) return false;
    mv.from = static_cast<int8_t>(from); mv.to = static_cast<int8_t>(to);
    mv.piece = static_cast<int8_t>(type_of(board[from]));
    mv.promotes = s.size() >= 5 && s[4] == '+';
    mv.captured = static_cast<int8_t>(type_of(board[to]));
    return true;
  }




  // gives_check: 1 / 0 when the caller already knows whether the move checks (a search tree does), -1 to find out.
  void do_move(const Move& mv, int gives_check = -1) {
    Move m = mv;
    if (!m.is_drop()) m.captured = static_cast<int8_t>(type_of(board[m.to]));   // whatever the caller knew: the board decides
    const uint64_t h = hash_after(last_entry().hash, m);
    apply(m);
    ++ply;
    push_history(h, gives_check < 0 ? in_check(side) : gives_check != 0, m);
  }
  void undo_move(const Move& m) {
    hist.pop_back();
    --ply;
    revert(m);
  }

  using Board::generate_moves;
  void generate_moves(std::vector<Move>& out) {
    MoveList l;
    generate_moves(l);
    out.assign(l.begin(), l.end());
  }
  std::vector<Move> generate_moves() { std::vector<Move> v; generate_moves(v); return v; }

  // ---- repetition (tests/test_repetition.py) -----------------------------------------------------
  // The position has now occurred four times -> repetition.  If every move one side made since the first
  // of those occurrences gave check, that side is the perpetual checker.
  void is_repetition(bool& rep, bool& my_check, bool& op_check) const {
    rep = my_check = op_check = false;
    const int last = history_size() - 1;
    if (entry(last).snap.rep_count < 4) return;
    const uint64_t h = entry(last).hash;
    int first = last;
    for (int i = last; i >= 0; i -= 2)
      if (entry(i).hash == h) first = i;
    rep = true;
    // moves leading to positions first+1 .. last; the move to position i was made by the side NOT to move at i
    bool mine = true, theirs = true;   // "mine" = side to move now
    bool any_mine = false, any_theirs = false;
    for (int i = first + 1; i <= last; ++i) {
      const bool mover_is_me = ((last - i) & 1) == 1;  // position `last` has me to move; the move into it was theirs
      if (mover_is_me) { any_mine = true; mine = mine && entry(i).gave_check; }
      else { any_theirs = true; theirs = theirs && entry(i).gave_check; }
    }
    my_check = any_mine && mine;
    op_check = any_theirs && theirs;
  }
  int repetition_count() const { return last_entry().snap.rep_count; }


  uint64_t hash() const { return last_entry().hash; }


 private:
  static PieceType letter_type(char ch) {
    switch (ch) {
      case 'K': case 'k': return KING; case 'G': case 'g': return GOLD; case 'S': case 's': return SILVER;
      case 'B': case 'b': return BISHOP; case 'R': case 'r': return ROOK; case 'P': case 'p': return PAWN;
      default: return EMPTY;
    }
  }

  void copy_board_from(const Position& o) { static_cast<Board&>(*this) = static_cast<const Board&>(o); }



  void push_history(uint64_t h, bool check, const Move& m) {
    hist.emplace_back();
    HistoryEntry& e = hist.back();
    e.hash = h;
    e.gave_check = check ? 1 : 0;
    e.move = m;
    for (int c = 0; c < 2; ++c)
      for (int t = 1; t < PIECE_TYPES; ++t) e.snap.pieces[c][t - 1] = bb[c][t];
    memcpy(e.snap.hand, hand, sizeof(hand));
    int c = 1;                                          // occurrences of this position so far (same side to move)
    for (int i = history_size() - 3; i >= 0; i -= 2)
      if (entry(i).hash == h) { c += entry(i).snap.rep_count; break; }   // the latest earlier occurrence knows the rest
    e.snap.rep_count = static_cast<uint8_t>(c > 255 ? 255 : c);
  }



};

// ---- network input (restated; see header comment) -------------------------------------------------
// 266 planes from the side to move's perspective (board rotated 180 degrees for WHITE):
//   8 history steps (current position first) x 33 planes:
//     0-9   own K,G,S,B,R,P,+S,+B,+R,+P      10-19 the opponent's
//     20-22 position seen once / twice / three or more times (constant planes)
//     23-27 own hand G,S,B,R,P as count/2     28-32 the opponent's hand
//   264 side to move (all ones when WHITE)    265 ply / 512
// Every plane is an occupancy plane or a constant, so the packed form (one bit per square + one value per
// plane, erweitern_55_b200/packing.py) is exact.
constexpr int kInputPlanes = 266, kHistory = 8, kPlanesPerStep = 33;

MS_HD inline Bitboard rotate180(Bitboard b) {   // square s -> 24 - s
  b = ((b >> 1) & 0x55555555u) | ((b & 0x55555555u) << 1);
  b = ((b >> 2) & 0x33333333u) | ((b & 0x33333333u) << 2);
  b = ((b >> 4) & 0x0F0F0F0Fu) | ((b & 0x0F0F0F0Fu) << 4);
#ifdef __CUDA_ARCH__
  return __byte_perm(b, 0, 0x0123) >> 7;
#else
  return __builtin_bswap32(b) >> 7;
#endif
}

inline void encode_packed(const Position& pos, uint32_t* bits, float* scale) {
  memset(bits, 0, kInputPlanes * sizeof(uint32_t));
  for (int c = 0; c < kInputPlanes; ++c) scale[c] = 1.f;
  const int me = pos.side;
  const int last = pos.history_size() - 1;
  for (int t = 0; t < kHistory; ++t) {
    const int idx = last - t;
    if (idx < 0) break;
    const Snapshot& sn = pos.entry(idx).snap;
    uint32_t* b = bits + t * kPlanesPerStep;
    float* sc = scale + t * kPlanesPerStep;
    if (me == BLACK) {
      for (int k = 0; k < 10; ++k) { b[k] = sn.pieces[BLACK][k]; b[10 + k] = sn.pieces[WHITE][k]; }
    } else {
      for (int k = 0; k < 10; ++k) { b[k] = rotate180(sn.pieces[WHITE][k]); b[10 + k] = rotate180(sn.pieces[BLACK][k]); }
    }
    const int rc = sn.rep_count >= 3 ? 3 : sn.rep_count;
    if (rc >= 1) b[20 + rc - 1] = kAllSquares;
    for (int k = 0; k < kHandTypes; ++k) {
      if (sn.hand[me][k]) { b[23 + k] = kAllSquares; sc[23 + k] = sn.hand[me][k] / 2.f; }
      if (sn.hand[me ^ 1][k]) { b[28 + k] = kAllSquares; sc[28 + k] = sn.hand[me ^ 1][k] / 2.f; }
    }
  }
  if (me == WHITE) bits[264] = kAllSquares;
  if (pos.ply > 0) { bits[265] = kAllSquares; scale[265] = static_cast<float>(pos.ply > 512 ? 512 : pos.ply) / 512.f; }
}

// ---- policy index (restated): 69 move planes x 25 squares, side to move's perspective --------------
//   planes 0-31: direction d (N, NE, E, SE, S, SW, W, NW) x distance 1..4, indexed by the ORIGIN square
//   planes 32-63: the same with promotion      planes 64-68: drop of G,S,B,R,P indexed by the DESTINATION
MS_HD inline int policy_index(const Move& m, int side) {
  auto persp = [&](int sq) { return side == BLACK ? sq : kSquares - 1 - sq; };
  if (m.is_drop()) return (64 + hand_index(static_cast<PieceType>(m.piece))) * kSquares + persp(m.to);
  const int from = persp(m.from), to = persp(m.to);
  const int dr = rank_of(to) - rank_of(from), dc = col_of(to) - col_of(from);
  const int adr = dr < 0 ? -dr : dr, adc = dc < 0 ? -dc : dc;
  const int dist = adr > adc ? adr : adc;
  const int sr = (dr > 0) - (dr < 0), sc = (dc > 0) - (dc < 0);
  // direction index of (sr, sc) in the order N, NE, E, SE, S, SW, W, NW (kDirR / kDirC)
  const int d = sr < 0 ? (sc == 0 ? 0 : (sc > 0 ? 1 : 7)) : (sr == 0 ? (sc > 0 ? 2 : 6) : (sc > 0 ? 3 : (sc == 0 ? 4 : 5)));
  return ((m.promotes ? 32 : 0) + d * 4 + (dist - 1)) * kSquares + from;
}

}  // namespace mshogi
