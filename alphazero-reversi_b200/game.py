"""Drop-in mirrors of the reference's game classes, executing on the GPU through the C ABI.

    reference                                   here
    src/game/board.py:10-431   Board        ->  Board
    src/game/game.py:9-192     ReversiGame  ->  ReversiGame

Same constructor signatures, attributes, return types and error behaviour (illegal moves return
False and never raise, game.py:47-48,70).  Every rule evaluation (legal mask, flips, auto-pass,
terminal/winner, canonical planes) is a batch-of-1 call of the CUDA kernels; there is no host
implementation of the rules in this package.  Batch users should call board_ops / Engine directly.
"""
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import _lib as L
from . import board_ops as ops


def _mask_to_moves(mask: int) -> List[Tuple[int, int]]:
    out = []
    m = int(mask)
    while m:
        i = (m & -m).bit_length() - 1
        m &= m - 1
        out.append(divmod(i, 8))  # ascending bit order (board.py:127-131)
    return out


class Board:
    """Bitboard Reversi board (reference: src/game/board.py:10-431)."""

    SIZE = 8
    BOARD_SIZE = 64
    EMPTY, BLACK, WHITE = 0, 1, 2
    RULES = L.RULES_REF

    def __init__(self, size: int = 8):
        if size != 8:
            raise ValueError("Only 8x8 board is supported")  # board.py:27-28
        self.size = size
        self.black = 0x0000000810000000
        self.white = 0x0000001008000000
        self.current_player = self.BLACK
        self.game_over = False
        self.winner = None
        self.move_history = []
        self.passed_moves_in_a_row = 0

    # -- helpers -------------------------------------------------------------------------
    def _arrays(self, player):
        return (np.array([self.black], dtype=np.uint64), np.array([self.white], dtype=np.uint64),
                np.array([player], dtype=np.uint8))

    def _legal_mask(self, player: int) -> int:
        b, w, s = self._arrays(player)
        return int(ops.legal_masks(b, w, s, rules=self.RULES)[0])

    @property
    def _board(self) -> np.ndarray:
        bits_b = np.unpackbits(np.array([self.black], dtype="<u8").view(np.uint8), bitorder="little")
        bits_w = np.unpackbits(np.array([self.white], dtype="<u8").view(np.uint8), bitorder="little")
        return (bits_b.astype(int) + 2 * bits_w.astype(int)).reshape(8, 8)

    def _ensure_board_updated(self) -> None:
        pass

    def _update_board_state(self) -> None:
        pass

    # -- reference API -------------------------------------------------------------------
    def copy(self) -> "Board":
        nb = Board(self.size)
        nb.black, nb.white = self.black, self.white
        nb.current_player = self.current_player
        nb.game_over, nb.winner = self.game_over, self.winner
        nb.move_history = self.move_history.copy()
        nb.passed_moves_in_a_row = self.passed_moves_in_a_row
        return nb

    def get_valid_moves(self, player: int = None) -> List[Tuple[int, int]]:
        if player is None:
            player = self.current_player
        return _mask_to_moves(self._legal_mask(player))

    def make_move(self, row: int, col: int, player: int = None) -> bool:
        if player is None:
            player = self.current_player
        if row == -1 and col == -1:  # explicit pass (board.py:151-167)
            if self._legal_mask(player):
                return False
            self.passed_moves_in_a_row += 1
            self.move_history.append((row, col, player))
            self.current_player = 3 - player
            if self.passed_moves_in_a_row >= 2:
                self.game_over = True
                self._determine_winner()
            return True
        row, col = int(row), int(col)
        if not (0 <= row < 8 and 0 <= col < 8):
            return False
        b, w, s = self._arrays(player)
        f = np.zeros(1, dtype=np.uint8)
        mv = np.array([row * 8 + col], dtype=np.uint8)
        ok, _ = ops.apply_moves(b, w, s, f, mv, rules=self.RULES, want_legal=False)
        if not ok[0]:
            return False
        self.black, self.white = int(b[0]), int(w[0])
        self.move_history.append((row, col, player))
        self.current_player = int(s[0])
        fl = int(f[0])
        self.passed_moves_in_a_row = 1 if fl & L.FLAG_PASSED else 0
        if fl & L.FLAG_OVER:
            self.game_over = True
            self.winner = (fl & L.FLAG_WINNER_MASK) >> L.FLAG_WINNER_SHIFT
        return True

    def is_valid_move(self, row: int, col: int, player: int = None) -> bool:
        """true-rules check of the reference (board.py:253-285): uses flips, not the legal mask"""
        if player is None:
            player = self.current_player
        if row < 0 or row >= 8 or col < 0 or col >= 8 or ((self.black | self.white) >> (row * 8 + col)) & 1:
            return False
        b, w, s = self._arrays(player)
        mv = np.array([row * 8 + col], dtype=np.uint8)
        return int(ops.flip_masks(b, w, s, mv, rules=L.RULES_STRICT)[0]) != 0

    def has_any_valid_move(self, player: int = None) -> bool:
        return len(self.get_valid_moves(player)) > 0

    def _get_flipped_pieces(self, move: Tuple[int, int], player: int):
        b, w, s = self._arrays(player)
        mv = np.array([move[0] * 8 + move[1]], dtype=np.uint8)
        return _mask_to_moves(int(ops.flip_masks(b, w, s, mv, rules=self.RULES)[0]))

    def _check_game_over(self) -> bool:
        if self.game_over:
            return True
        if not self.has_any_valid_move(self.BLACK) and not self.has_any_valid_move(self.WHITE):
            self.game_over = True
            self._determine_winner()
            return True
        return False

    def _determine_winner(self) -> None:
        bc, wc = self.get_score()
        self.winner = self.BLACK if bc > wc else self.WHITE if wc > bc else 0

    def __call__(self, row: int, col: int, player: int = None) -> bool:
        return self.make_move(row, col, player)

    def get_board_state(self) -> np.ndarray:
        return self._board.copy()

    def get_score(self) -> Tuple[int, int]:
        return (self.bit_count(self.black), self.bit_count(self.white))

    @staticmethod
    def bit_count(x: int) -> int:
        return bin(int(x) & 0xFFFFFFFFFFFFFFFF).count("1")

    def __str__(self) -> str:
        symbols = {0: ".", 1: "B", 2: "W"}
        bd = self._board
        rows = [" ".join(symbols[int(bd[i, j])] for j in range(8)) for i in range(8)]
        status = ["\n".join(rows)]
        status.append(f"Current player: {'Black' if self.current_player == self.BLACK else 'White'}")
        bc, wc = self.get_score()
        status.append(f"Score - Black: {bc}, White: {wc}")
        if self.game_over:
            if self.winner == 0:
                status.append("Game over! It's a draw!")
            else:
                status.append(f"Game over! {'Black' if self.winner == self.BLACK else 'White'} wins!")
        return "\n".join(status)


class ReversiGame:
    """Game wrapper (reference: src/game/game.py:9-192)."""

    def __init__(self, size: int = 8):
        self.board = Board(size)
        self.size = size
        self.current_player = Board.BLACK
        self.game_over = False
        self.winner = None
        self.move_history = []

    def reset(self) -> None:
        self.__init__(self.size)

    def make_move(self, row: int, col: int) -> bool:
        if self.game_over:
            return False
        board_before = self.board.copy()
        move_made = self.board.make_move(row, col, self.current_player)
        if move_made:
            self.move_history.append({"player": self.current_player, "move": (row, col),
                                      "board_before": board_before, "board_after": self.board.copy()})
            self.game_over = self.board.game_over
            self.winner = self.board.winner
            self.current_player = self.board.current_player
        return move_made

    def get_valid_moves(self) -> List[Tuple[int, int]]:
        return self.board.get_valid_moves(self.current_player)

    def is_game_over(self) -> bool:
        return self.board.game_over

    def get_winner(self) -> Optional[int]:
        return self.board.winner if self.game_over else None

    def get_score(self) -> Tuple[int, int]:
        return self.board.get_score()

    def get_board_state(self) -> np.ndarray:
        return self.board.get_board_state()

    def get_current_player(self) -> int:
        return self.current_player

    def get_move_history(self) -> List[Dict[str, Any]]:
        return self.move_history.copy()

    def get_canonical_state(self) -> np.ndarray:
        """(3,8,8) float32: own discs, opponent discs, legal mask (game.py:131-162)"""
        b = np.array([self.board.black], dtype=np.uint64)
        w = np.array([self.board.white], dtype=np.uint64)
        s = np.array([self.current_player], dtype=np.uint8)
        return ops.encode_planes(b, w, s, L.PLANES_F32_NCHW, rules=self.board.RULES)[0]

    def copy(self) -> "ReversiGame":
        g = ReversiGame(self.size)
        g.board = self.board.copy()
        g.current_player = self.current_player
        g.game_over = self.game_over
        g.winner = self.winner
        g.move_history = self.move_history.copy()
        return g

    def __str__(self) -> str:
        return str(self.board)
