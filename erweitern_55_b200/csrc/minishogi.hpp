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
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace mshogi {

constexpr int kN = 5, kSquares = 25;
enum Color : int { BLACK = 0, WHITE = 1 };
enum PieceType : int { EMPTY = 0, KING, GOLD, SILVER, BISHOP, ROOK, PAWN, PRO_SILVER, HORSE, DRAGON, TOKIN, PIECE_TYPES };
constexpr int kHandTypes = 5;  // GOLD, SILVER, BISHOP, ROOK, PAWN
constexpr PieceType kHandPiece[kHandTypes] = {GOLD, SILVER, BISHOP, ROOK, PAWN};

inline int hand_index(PieceType t) {  // unpromoted type -> hand slot
  switch (t) { case GOLD: return 0; case SILVER: return 1; case BISHOP: return 2; case ROOK: return 3; case PAWN: return 4; default: return -1; }
}
inline PieceType promote(PieceType t) {
  switch (t) { case SILVER: return PRO_SILVER; case BISHOP: return HORSE; case ROOK: return DRAGON; case PAWN: return TOKIN; default: return t; }
}
inline PieceType unpromote(PieceType t) {
  switch (t) { case PRO_SILVER: return SILVER; case HORSE: return BISHOP; case DRAGON: return ROOK; case TOKIN: return PAWN; default: return t; }
}
inline bool can_promote(PieceType t) { return t == SILVER || t == BISHOP || t == ROOK || t == PAWN; }

// square = rank * 5 + col; rank 0 = 'a' (White's home row), col 0 = file 5 (left edge as SFEN prints it)
inline int sq_of(int rank, int col) { return rank * kN + col; }
inline int rank_of(int sq) { return sq / kN; }
inline int col_of(int sq) { return sq % kN; }

struct Move {
  int8_t from = -1;      // -1: drop
  int8_t to = 0;
  int8_t piece = 0;      // PieceType moved (before promotion) or dropped
  int8_t promotes = 0;
  int8_t captured = 0;   // PieceType on `to` (as it stood, promoted or not), filled by generate/do
  bool is_drop() const { return from < 0; }
  bool operator==(const Move& o) const { return from == o.from && to == o.to && piece == o.piece && promotes == o.promotes; }
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

inline uint32_t step_mask(PieceType t) {  // bit d set: may step one square in direction d (BLACK orientation)
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
inline uint32_t slide_mask(PieceType t) {
  switch (t) {
    case BISHOP: case HORSE: return (1u << 1) | (1u << 3) | (1u << 5) | (1u << 7);
    case ROOK: case DRAGON: return (1u << 0) | (1u << 2) | (1u << 4) | (1u << 6);
    default: return 0;
  }
}

struct MoveList {                    // no position has anywhere near 256 legal moves on a 5x5 board
  Move m[256];
  int n = 0;
  void push(const Move& mv) { if (n < 256) m[n++] = mv; }
  const Move* begin() const { return m; }
  const Move* end() const { return m + n; }
  bool empty() const { return n == 0; }
  int size() const { return n; }
};

struct Snapshot {                    // what the input encoder needs of a past position
  int8_t board[kSquares];           // type | color << 4
  int8_t hand[2][kHandTypes];
  uint8_t rep_count;                // occurrences of that position up to then (1..)
};

struct HistoryEntry {                // one per position of the game so far (index 0 = the position set_sfen gave)
  uint64_t hash;
  Snapshot snap;
  uint8_t gave_check;               // the move leading here gave check (index 0: unused)
  Move move;                        // the move leading here (index 0: unused)
};

class Position {
 public:
  int8_t board[kSquares];            // 0 = empty, else type | (color << 4)
  int8_t hand[2][kHandTypes];
  int side = BLACK;
  int ply = 0;
  int8_t king_sq[2] = {-1, -1};      // cached king squares (-1: no king, as in some test positions)
  std::vector<HistoryEntry> hist;    // history for repetition / encoding / tree reuse; hist[ply] = current position

  Position() { set_start_position(); }

  static int type_of(int8_t p) { return p & 15; }
  static int color_of(int8_t p) { return (p >> 4) & 1; }
  static int8_t make_piece(int type, int color) { return static_cast<int8_t>(type | (color << 4)); }

  void clear() {
    memset(board, 0, sizeof(board));
    memset(hand, 0, sizeof(hand));
    side = BLACK; ply = 0;
    king_sq[0] = king_sq[1] = -1;
    hist.clear();
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
        board[sq_of(r, c)] = make_piece(t, color);
        if (t == KING) king_sq[color] = static_cast<int8_t>(sq_of(r, c));
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
    push_history(false, Move());
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

  // ---- attacks -------------------------------------------------------------------------------
  bool attacked_by(int sq, int by) const {
    const int r = rank_of(sq), c = col_of(sq);
    const int sgn = by == BLACK ? 1 : -1;  // BLACK's forward is -rank
    for (int d = 0; d < 8; ++d) {
      // a piece of `by` standing at sq - dir attacks sq if it can step/slide in dir
      const int dr = kDirR[d] * sgn, dc = kDirC[d];
      int rr = r - dr, cc = c - dc;
      if (rr < 0 || rr >= kN || cc < 0 || cc >= kN) continue;
      int8_t p = board[sq_of(rr, cc)];
      if (p && color_of(p) == by && ((step_mask(static_cast<PieceType>(type_of(p))) | slide_mask(static_cast<PieceType>(type_of(p)))) >> d & 1)) return true;
      // sliders further away
      while (!p) {
        rr -= dr; cc -= dc;
        if (rr < 0 || rr >= kN || cc < 0 || cc >= kN) break;
        p = board[sq_of(rr, cc)];
        if (p) { if (color_of(p) == by && (slide_mask(static_cast<PieceType>(type_of(p))) >> d & 1)) return true; break; }
      }
    }
    return false;
  }
  int king_square(int color) const { return king_sq[color]; }
  bool in_check(int color) const {
    const int k = king_square(color);
    return k >= 0 && attacked_by(k, color ^ 1);
  }

  // ---- make / unmake (board + hands + side only; history handled by do_move / undo_move) -------------
  void apply(const Move& m) {
    const int us = side;
    if (m.is_drop()) {
      board[m.to] = make_piece(m.piece, us);
      --hand[us][hand_index(static_cast<PieceType>(m.piece))];
    } else {
      if (m.captured) ++hand[us][hand_index(unpromote(static_cast<PieceType>(m.captured)))];
      board[m.to] = make_piece(m.promotes ? promote(static_cast<PieceType>(m.piece)) : m.piece, us);
      board[m.from] = 0;
      if (m.piece == KING) king_sq[us] = m.to;
    }
    side ^= 1;
  }
  void revert(const Move& m) {
    side ^= 1;
    const int us = side;
    if (m.is_drop()) {
      board[m.to] = 0;
      ++hand[us][hand_index(static_cast<PieceType>(m.piece))];
    } else {
      board[m.from] = make_piece(m.piece, us);
      if (m.piece == KING) king_sq[us] = m.from;
      board[m.to] = m.captured ? make_piece(m.captured, us ^ 1) : 0;
      if (m.captured) --hand[us][hand_index(unpromote(static_cast<PieceType>(m.captured)))];
    }
  }

  void do_move(const Move& m) {
    apply(m);
    ++ply;
    push_history(in_check(side), m);
  }
  void undo_move(const Move& m) {
    hist.pop_back();
    --ply;
    revert(m);
  }

  // ---- move generation ------------------------------------------------------------------------
  // Fully legal moves of the side to move (own king never left in check; nifu, last-rank pawn and
  // pawn-drop-mate excluded).
  void generate_moves(MoveList& out) {
    out.n = 0;
    MoveList pseudo;
    pseudo_moves(pseudo);
    const int us = side;
    const int ksq = king_sq[us];
    const bool checked = ksq >= 0 && attacked_by(ksq, us ^ 1);
    const int opp_k = king_sq[us ^ 1];
    for (const Move& m : pseudo) {
      // Playing the move and testing for check is only needed when it could expose the king: we are in
      // check, the king itself moves, or the piece leaves a line through the king (possible pin).  A pawn
      // drop is additionally played out when it checks the enemy king (pawn-drop mate is illegal).
      bool test = checked || ksq < 0 || m.piece == KING;
      if (!test && !m.is_drop()) {
        const int dr = rank_of(m.from) - rank_of(ksq), dc = col_of(m.from) - col_of(ksq);
        test = dr == 0 || dc == 0 || dr == dc || dr == -dc;
      }
      const bool pawn_drop_check = m.is_drop() && m.piece == PAWN && opp_k >= 0 && col_of(opp_k) == col_of(m.to) &&
                                   rank_of(opp_k) == rank_of(m.to) + (us == BLACK ? -1 : 1);
      if (!test && !pawn_drop_check) { out.push(m); continue; }
      apply(m);
      bool ok = !in_check(us);
      if (ok && pawn_drop_check && !has_legal_move()) ok = false;  // uchifuzume
      revert(m);
      if (ok) out.push(m);
    }
  }
  void generate_moves(std::vector<Move>& out) {
    MoveList l;
    generate_moves(l);
    out.assign(l.begin(), l.end());
  }
  std::vector<Move> generate_moves() { std::vector<Move> v; generate_moves(v); return v; }

  bool has_legal_move() {
    MoveList pseudo;
    pseudo_moves(pseudo);
    const int us = side;
    for (const Move& m : pseudo) {
      apply(m);
      bool ok = !in_check(us);
      if (ok && m.is_drop() && m.piece == PAWN && in_check(side) && !has_legal_move()) ok = false;
      revert(m);
      if (ok) return true;
    }
    return false;
  }

  // ---- repetition (tests/test_repetition.py) -----------------------------------------------------
  // The position has now occurred four times -> repetition.  If every move one side made since the first
  // of those occurrences gave check, that side is the perpetual checker.
  void is_repetition(bool& rep, bool& my_check, bool& op_check) const {
    rep = my_check = op_check = false;
    const int last = static_cast<int>(hist.size()) - 1;
    if (hist[last].snap.rep_count < 4) return;
    const uint64_t h = hist[last].hash;
    int count = 0, first = last;
    for (int i = last; i >= 0; i -= 2)
      if (hist[i].hash == h) { ++count; first = i; }
    if (count < 4) return;
    rep = true;
    // moves leading to positions first+1 .. last; the move to position i was made by the side NOT to move at i
    bool mine = true, theirs = true;   // "mine" = side to move now
    bool any_mine = false, any_theirs = false;
    for (int i = first + 1; i <= last; ++i) {
      const bool mover_is_me = ((last - i) & 1) == 1;  // position `last` has me to move; the move into it was theirs
      if (mover_is_me) { any_mine = true; mine = mine && hist[i].gave_check; }
      else { any_theirs = true; theirs = theirs && hist[i].gave_check; }
    }
    my_check = any_mine && mine;
    op_check = any_theirs && theirs;
  }
  int repetition_count() const { return hist.back().snap.rep_count; }

  // ---- mate search (tests/test_checkmate.py) ----------------------------------------------------
  // Depth-limited AND/OR search over checking sequences: the side to move must check on every move.
  bool solve_checkmate_dfs(int depth, Move& best) {
    return depth > 0 && mate_attack(depth, &best);
  }

  uint64_t hash() const { return hist.back().hash; }

 private:
  static PieceType letter_type(char ch) {
    switch (ch) {
      case 'K': case 'k': return KING; case 'G': case 'g': return GOLD; case 'S': case 's': return SILVER;
      case 'B': case 'b': return BISHOP; case 'R': case 'r': return ROOK; case 'P': case 'p': return PAWN;
      default: return EMPTY;
    }
  }

  static uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
  }
  uint64_t compute_hash() const {
    uint64_t h = side == BLACK ? 0x1234567ull : 0x7654321ull;
    for (int s = 0; s < kSquares; ++s)
      if (board[s]) h ^= mix(static_cast<uint64_t>(s) * 64 + static_cast<uint8_t>(board[s]));
    for (int c = 0; c < 2; ++c)
      for (int k = 0; k < kHandTypes; ++k) h ^= mix(4096 + (c * kHandTypes + k) * 8 + hand[c][k]);
    return h;
  }
  void push_history(bool check, const Move& m) {
    HistoryEntry e;
    e.hash = compute_hash();
    e.gave_check = check ? 1 : 0;
    e.move = m;
    memcpy(e.snap.board, board, sizeof(board));
    memcpy(e.snap.hand, hand, sizeof(hand));
    int c = 1;                                          // occurrences of this position so far (same side to move)
    for (int i = static_cast<int>(hist.size()) - 2; i >= 0; i -= 2)
      if (hist[i].hash == e.hash) { c += hist[i].snap.rep_count; break; }   // the latest earlier occurrence knows the rest
    e.snap.rep_count = static_cast<uint8_t>(c > 255 ? 255 : c);
    hist.push_back(e);
  }

  void pseudo_moves(MoveList& out) const {
    const int us = side;
    const int sgn = us == BLACK ? 1 : -1;
    const int far = us == BLACK ? 0 : kN - 1;  // promotion rank
    for (int s = 0; s < kSquares; ++s) {
      const int8_t p = board[s];
      if (!p || color_of(p) != us) continue;
      const PieceType t = static_cast<PieceType>(type_of(p));
      const uint32_t steps = step_mask(t), slides = slide_mask(t);
      for (int d = 0; d < 8; ++d) {
        const bool slide = slides >> d & 1;
        if (!slide && !(steps >> d & 1)) continue;
        const int dr = kDirR[d] * sgn, dc = kDirC[d];
        int rr = rank_of(s) + dr, cc = col_of(s) + dc;
        while (rr >= 0 && rr < kN && cc >= 0 && cc < kN) {
          const int to = sq_of(rr, cc);
          const int8_t q = board[to];
          if (q && color_of(q) == us) break;
          Move m;
          m.from = static_cast<int8_t>(s); m.to = static_cast<int8_t>(to); m.piece = static_cast<int8_t>(t);
          m.captured = static_cast<int8_t>(type_of(q));
          const bool zone = rank_of(s) == far || rr == far;
          if (can_promote(t) && zone) { Move pm = m; pm.promotes = 1; out.push(pm); }
          if (!(t == PAWN && rr == far)) out.push(m);  // a pawn on the last rank must promote
          if (q || !slide) break;
          rr += dr; cc += dc;
        }
      }
    }
    for (int k = 0; k < kHandTypes; ++k) {
      if (!hand[us][k]) continue;
      const PieceType t = kHandPiece[k];
      for (int to = 0; to < kSquares; ++to) {
        if (board[to]) continue;
        if (t == PAWN) {
          if (rank_of(to) == far) continue;
          bool nifu = false;
          for (int r = 0; r < kN; ++r)
            if (board[sq_of(r, col_of(to))] == make_piece(PAWN, us)) nifu = true;
          if (nifu) continue;
        }
        Move m;
        m.from = -1; m.to = static_cast<int8_t>(to); m.piece = static_cast<int8_t>(t);
        out.push(m);
      }
    }
  }

  bool mate_attack(int depth, Move* best) {
    MoveList moves;
    generate_moves(moves);
    for (const Move& m : moves) {
      apply(m);
      bool mated = false;
      if (in_check(side)) mated = mate_defend(depth - 1);
      revert(m);
      if (mated) { if (best) *best = m; return true; }
    }
    return false;
  }
  // true iff the side to move (in check) is mated within `depth` further plies whatever it plays
  bool mate_defend(int depth) {
    MoveList moves;
    generate_moves(moves);
    if (moves.empty()) return true;
    if (depth <= 1) return false;
    for (const Move& m : moves) {
      apply(m);
      const bool still = mate_attack(depth - 1, nullptr);
      revert(m);
      if (!still) return false;
    }
    return true;
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

inline void encode_packed(const Position& pos, uint32_t* bits, float* scale) {
  for (int c = 0; c < kInputPlanes; ++c) { bits[c] = 0u; scale[c] = 1.f; }
  const int me = pos.side;
  const uint32_t all = (1u << kSquares) - 1u;
  const int last = static_cast<int>(pos.hist.size()) - 1;
  for (int t = 0; t < kHistory; ++t) {
    const int idx = last - t;
    if (idx < 0) break;
    const Snapshot& sn = pos.hist[idx].snap;
    const int base = t * kPlanesPerStep;
    for (int s = 0; s < kSquares; ++s) {
      const int8_t p = sn.board[s];
      if (!p) continue;
      const int sq = me == BLACK ? s : kSquares - 1 - s;
      const int type = Position::type_of(p), color = Position::color_of(p);
      bits[base + (color == me ? 0 : 10) + (type - 1)] |= 1u << sq;
    }
    const int rc = sn.rep_count >= 3 ? 3 : sn.rep_count;
    if (rc >= 1) bits[base + 20 + rc - 1] = all;
    for (int k = 0; k < kHandTypes; ++k) {
      if (sn.hand[me][k]) { bits[base + 23 + k] = all; scale[base + 23 + k] = sn.hand[me][k] / 2.f; }
      if (sn.hand[me ^ 1][k]) { bits[base + 28 + k] = all; scale[base + 28 + k] = sn.hand[me ^ 1][k] / 2.f; }
    }
  }
  if (me == WHITE) bits[264] = all;
  if (pos.ply > 0) { bits[265] = all; scale[265] = static_cast<float>(pos.ply > 512 ? 512 : pos.ply) / 512.f; }
}

// ---- policy index (restated): 69 move planes x 25 squares, side to move's perspective --------------
//   planes 0-31: direction d (N, NE, E, SE, S, SW, W, NW) x distance 1..4, indexed by the ORIGIN square
//   planes 32-63: the same with promotion      planes 64-68: drop of G,S,B,R,P indexed by the DESTINATION
inline int policy_index(const Move& m, int side) {
  auto persp = [&](int sq) { return side == BLACK ? sq : kSquares - 1 - sq; };
  if (m.is_drop()) return (64 + hand_index(static_cast<PieceType>(m.piece))) * kSquares + persp(m.to);
  const int from = persp(m.from), to = persp(m.to);
  const int dr = rank_of(to) - rank_of(from), dc = col_of(to) - col_of(from);
  const int dist = std::max(dr < 0 ? -dr : dr, dc < 0 ? -dc : dc);
  const int sr = (dr > 0) - (dr < 0), sc = (dc > 0) - (dc < 0);
  int d = 0;
  for (int k = 0; k < 8; ++k)
    if (kDirR[k] == sr && kDirC[k] == sc) d = k;
  return ((m.promotes ? 32 : 0) + d * 4 + (dist - 1)) * kSquares + from;
}

}  // namespace mshogi
