"""Predicate programs: queries -> the binary blob `fei_scan_*` takes (include/feiscan_prog.h).

A program holds up to 32 queries; a query is the AND of conditions.  Every string-valued
condition becomes one output bit of the multi-output byte DFA of the field it reads
(fei_b200.regexc), so all conditions of all queries on one field cost one pass over it.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .regexc import Dfa, Pattern, compile_patterns

MAGIC, VERSION = 0x50494546, 1
C_CONST, C_BODY, C_SLOT, C_FLAGS, C_NAME, C_DATE_CMP, C_FOLDER_SET, C_STATUS_SET, C_RECBITS, C_TS_CMP = range(10)
CMP = {">": 0, "<": 1, ">=": 2, "<=": 3, "=": 4, "!=": 5}
NAME_FILENAME, NAME_ID, NAME_HOST, NAME_TS_STR, NAME_DATE_STR = 0, 1, 2, 3, 4     # 3 / 4: str(metadata["timestamp"]) / str(metadata["date"]), formatted by the kernels
MAX_AUX = 4                            # aux verdict columns per corpus (FEI_C_RECBITS)
SMEM_TABLE_LIMIT = 200 * 1024          # body automaton must fit the CTA's shared memory
HEAD_SMEM_WINDOW = 32 * 1024           # scan.cu kHeadProgSmem: the program head the head kernels stage per CTA
HEAD_DIRECT_LIMIT = 12 * 1024          # a head automaton up to this size is stored byte-indexed


@dataclass(frozen=True)
class Cond:
    """One condition.  kind + what it reads:
       const(value) | body(pattern) | slot(field, mode, empty_if_missing, pattern) | flags(pattern)
       | name(which, pattern) | date_cmp(op, micros) | folder_set(bits) | status_set(bits)
       | recbits(which = aux column, negate) | ts_cmp(op, i64)"""
    kind: int
    pattern: Optional[Pattern] = None
    negate: bool = False
    field: str = ""              # header field name (slot)
    mode: int = 0                # slot: 0 = case-insensitive first key, 1 = exact key, 2 = any key (OR over all header values)
    empty_if_missing: bool = False
    if_missing: int = 0          # slot: 0/1 result when absent, 2 = next condition is the fallback
    which: int = 0               # name field
    op: int = 0                  # date_cmp
    i64: int = 0
    set64: int = 0
    value: bool = False          # const


def const(v: bool) -> Cond:
    return Cond(C_CONST, value=bool(v))


def _pad16(b: bytearray) -> None:
    while len(b) % 16:
        b.append(0)


_DFA_CACHE: Dict[Tuple, Dfa] = {}


class _FieldDfa:
    """Patterns that read the same field, deduplicated -> output bits."""

    def __init__(self, what: str):
        self.what = what
        self.patterns: List[Pattern] = []
        self.index: Dict[Tuple, int] = {}

    def bit(self, p: Pattern) -> int:
        key = (p.kind, p.text, int(p.flags))
        b = self.index.get(key)
        if b is None:
            b = len(self.patterns)
            if b >= 32:
                raise NotImplementedError(f"more than 32 distinct patterns on {self.what} in one program")
            self.index[key] = b
            self.patterns.append(p)
        return b

    def compile(self, sticky: bool = False) -> Dfa:
        """Compiled automata are cached per process: repeated queries / the default filters skip the (host-side,
        Python) subset construction, which costs 10-900 ms for Unicode-class-heavy patterns."""
        sticky = sticky and len(self.patterns) == 1
        key = (tuple((p.kind, p.text, int(p.flags)) for p in self.patterns), sticky)
        d = _DFA_CACHE.get(key)
        if d is None:
            d = compile_patterns(self.patterns, sticky=sticky)
            if len(_DFA_CACHE) >= 256:
                _DFA_CACHE.pop(next(iter(_DFA_CACHE)))
            _DFA_CACHE[key] = d
        return d


def tile_byte_perm(b: np.ndarray) -> np.ndarray:
    """Byte substitution applied to the body tiles at pack time (fei_b200/csrc/corpus.h): bit 5 ^= bit 6."""
    b = b.astype(np.int64)
    return b ^ ((b >> 1) & 0x20)


def serialize_dfa(d: Dfa, blob: bytearray, direct_limit: int = 0, tile_bytes: bool = False) -> int:
    """Append descriptor + tables; returns the descriptor offset.  States are renumbered so that
    states with a non-empty `out` come first (the kernels test `state < n_acc`, and the body kernel
    records accepting state k as bit k without a table lookup)."""
    n = d.n_states
    if n > 65535:
        raise NotImplementedError("automaton has more than 65535 states")
    accepting = d.out != 0
    order = np.concatenate([np.nonzero(accepting)[0], np.nonzero(~accepting)[0]])
    new_id = np.empty(n, dtype=np.int64)
    new_id[order] = np.arange(n)
    n_acc = int(accepting.sum())
    direct = direct_limit and n * 516 + n * 8 + 256 <= direct_limit
    cls = d.cls.astype(np.uint8)
    if direct:
        trans = new_id[d.trans[order]].astype(np.uint16)              # [n, 256]
        ncols = 256
        if tile_bytes:                                                # columns addressed by the stored (permuted) byte
            trans = trans[:, tile_byte_perm(np.arange(256))]
    else:
        trans = new_id[d.ctrans[order]].astype(np.uint16)             # [n, ncls]
        ncols = trans.shape[1]
        if tile_bytes:
            cls = cls[tile_byte_perm(np.arange(256))]
    # pad rows so that (row_stride / 2) is odd: state s starts at bank (s * stride/2) % 32, which walks
    # all 32 banks instead of piling every row onto the same ones (bank-conflict spreading)
    stride = ncols + (ncols & 1)
    if (stride // 2) % 2 == 0:
        stride += 2
    padded = np.zeros((n, stride), dtype=np.uint16)
    padded[:, :ncols] = trans
    trans = padded
    out = d.out[order].astype(np.uint32)
    endout = d.endout[order].astype(np.uint32)
    start = int(new_id[d.start])
    _pad16(blob)
    desc_off = len(blob)
    blob.extend(b"\0" * 64)
    _pad16(blob)
    off_trans = len(blob)
    blob.extend(trans.tobytes()); _pad16(blob)
    trans_bytes = len(blob) - off_trans
    off_out = len(blob); blob.extend(out.tobytes()); _pad16(blob)
    off_endout = len(blob); blob.extend(endout.tobytes()); _pad16(blob)
    off_cls = len(blob); blob.extend(cls.tobytes()); _pad16(blob)
    table_bytes = len(blob) - off_trans
    empty_acc = int(out[start]) | int(endout[start])
    sticky = 0
    if d.sticky_state >= -1:
        sticky = 0xFFFFFFFF if d.sticky_state < 0 else 1 + int(new_id[d.sticky_state])
    struct.pack_into("<16I", blob, desc_off, n, ncols, start, n_acc, off_trans, trans_bytes, off_out, off_endout,
                     off_cls, d.n_patterns, empty_acc, table_bytes, stride, sticky, 0, 0)
    return desc_off


class ProgramBuilder:
    def __init__(self):
        self.queries: List[List[Cond]] = []

    def add_query(self, conds: Sequence[Cond]) -> int:
        if len(self.queries) >= 32:
            raise NotImplementedError("at most 32 queries per program")
        self.queries.append(list(conds) if conds else [const(True)])
        return len(self.queries) - 1

    def build(self) -> bytes:
        body = _FieldDfa("content")
        flags = _FieldDfa("flags")
        names = [_FieldDfa("filename"), _FieldDfa("id"), _FieldDfa("hostname"), _FieldDfa("str(timestamp)"), _FieldDfa("str(date)")]
        slots: List[Tuple[str, int, bool]] = []          # (field, mode, empty_if_missing)
        slot_dfas: List[_FieldDfa] = []
        slot_index: Dict[Tuple, int] = {}
        recs: List[Tuple] = []                           # packed cond tuples
        qranges: List[Tuple[int, int]] = []
        head_mask = body_mask = slot_mask = name_mask = 0
        for qi, conds in enumerate(self.queries):
            begin = len(recs)
            for c in conds:
                ref = bit = 0
                if c.kind == C_BODY:
                    bit = body.bit(c.pattern); body_mask |= 1 << qi
                else:
                    head_mask |= 1 << qi
                    if c.kind == C_SLOT:
                        slot_mask |= 1 << qi
                        key = (c.field if c.mode == 1 else c.field.lower(), c.mode, c.empty_if_missing)
                        ref = slot_index.get(key)
                        if ref is None:
                            ref = len(slots)
                            if ref >= 16:
                                raise NotImplementedError("more than 16 distinct header fields in one program")
                            slot_index[key] = ref
                            slots.append(key); slot_dfas.append(_FieldDfa(f"header {c.field}"))
                        bit = slot_dfas[ref].bit(c.pattern)
                    elif c.kind == C_FLAGS:
                        bit = flags.bit(c.pattern)
                    elif c.kind == C_NAME:
                        name_mask |= 1 << qi
                        ref = c.which; bit = names[c.which].bit(c.pattern)
                    elif c.kind == C_CONST:
                        bit = 1 if c.value else 0
                    elif c.kind == C_RECBITS:
                        if not 0 <= c.which < MAX_AUX:
                            raise NotImplementedError("aux column index out of range")
                        ref = c.which
                recs.append((c.kind, ref, bit, 1 if c.negate else 0, c.if_missing, c.op, c.i64, c.set64))
            qranges.append((begin, len(recs)))

        # small head automata are stored byte-indexed (one lookup per byte in the head kernels) as long as the whole
        # program head still fits the kernels' shared-memory window; otherwise class-indexed (two lookups, smaller)
        for head_direct in (HEAD_DIRECT_LIMIT, 0):
            blob = bytearray(96)
            _pad16(blob)
            off_conds = len(blob)
            for kind, ref, bit, neg, ifm, op, i64, set64 in recs:
                blob.extend(struct.pack("<6B2xqQQ", kind, ref, bit, neg, ifm, op, i64, set64 & 0xFFFFFFFFFFFFFFFF, 0))
            _pad16(blob)
            off_queries = len(blob)
            for a, b in qranges:
                blob.extend(struct.pack("<4I", a, b, 0, 0))
            _pad16(blob)
            off_slots = len(blob)
            blob.extend(b"\0" * (16 * len(slots)))
            off_key = 0
            if slots:
                key_pats = [Pattern("regex", "", 0) if mode == 2 else Pattern("exact_equals", f) if mode == 1 else Pattern("equals", f)
                            for f, mode, _e in slots]
                off_key = serialize_dfa(compile_patterns(key_pats), blob)          # only run over the key dictionary (k_key_lut)
                for si, (f, mode, empty) in enumerate(slots):
                    off_val = serialize_dfa(slot_dfas[si].compile(), blob, head_direct)
                    struct.pack_into("<4I", blob, off_slots + 16 * si, mode, off_val, 1 if empty else 0, 0)
            off_flags = serialize_dfa(flags.compile(), blob, head_direct) if flags.patterns else 0
            off_names = [serialize_dfa(nf.compile(), blob, head_direct) if nf.patterns else 0 for nf in names]
            _pad16(blob)
            if len(blob) <= HEAD_SMEM_WINDOW:
                break
        head_bytes = len(blob)               # the content automaton goes last: [0, head_bytes) is what the head kernels stage
        off_body = serialize_dfa(body.compile(sticky=True), blob, SMEM_TABLE_LIMIT, tile_bytes=True) if body.patterns else 0
        if off_body:
            tb = struct.unpack_from("<I", blob, off_body + 44)[0]
            if tb > 220 * 1024:
                raise NotImplementedError(f"content automaton needs {tb} bytes of shared memory (limit 220 KiB)")
        _pad16(blob)
        struct.pack_into("<22I", blob, 0, MAGIC, VERSION, len(blob), len(self.queries), len(recs), off_conds, off_queries,
                         len(slots), off_slots, off_key, off_body, off_flags, off_names[0], off_names[1], off_names[2],
                         head_mask, body_mask, slot_mask, name_mask, head_bytes, off_names[3], off_names[4])
        return bytes(blob)


def content_batch_program(patterns: Sequence[Pattern]) -> bytes:
    """One query per pattern, each a single `content` condition (BASELINE configs[2])."""
    pb = ProgramBuilder()
    for p in patterns:
        pb.add_query([Cond(C_BODY, pattern=p)])
    return pb.build()
