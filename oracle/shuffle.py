"""Oracle (test infrastructure): the reference's shuffle protocol with P logical ranks in one process.

Restates sdc/shuffle_utils.py:119-298 over the single-process transport's semantics
(sdc/transport/hpat_transport_single_process.cpp:106-119: alltoall / alltoallv are memcpy between the
send and receive buffers):
  1. every rank counts its rows per destination ............ update_shuffle_meta, shuffle_utils.py:119-140
  2. alltoall of the counts, displacements = exclusive sums  finalize_shuffle_meta :152-162
  3. rows are written to the send buffer at send_disp[dest] + tmp_offset[dest] (stable within a
     destination) and every column goes through one alltoallv  :229-270
  4. block partition of n rows over P ranks: ceil(n / P) rows each  sdc/_distributed.h:77-102
`dest_of` is the partition function: any pure function of the key is a valid choice (the reference
uses hash(key) % n_pes); tests pass the device partition's own destinations so that only the PROTOCOL is
compared.
"""
import numpy as np


def block_range(n_total, rank, P):
    per = -(-n_total // P)
    start = min(n_total, rank * per)
    return start, min(n_total, start + per)


def calc_disp(counts):
    disp = np.zeros(len(counts), dtype=np.int64)
    disp[1:] = np.cumsum(counts)[:-1]
    return disp


def shuffle(columns_per_rank, dests_per_rank):
    """columns_per_rank[r] = list of 1-D arrays held by rank r; dests_per_rank[r][i] = destination of row i.
    -> received[r] = list of arrays on rank r after the shuffle (rows grouped by source rank, stable)."""
    P = len(columns_per_rank)
    send_counts = np.zeros((P, P), dtype=np.int64)
    for r in range(P):
        for d in dests_per_rank[r]:
            send_counts[r, d] += 1
    recv_counts = send_counts.T.copy()                      # the alltoall of the counts
    received = []
    for r in range(P):
        ncols = len(columns_per_rank[r])
        outs = [np.empty(int(recv_counts[r].sum()), dtype=columns_per_rank[r][c].dtype) for c in range(ncols)]
        rdisp = calc_disp(recv_counts[r])
        for src in range(P):
            sel = np.flatnonzero(np.asarray(dests_per_rank[src]) == r)      # stable within a destination
            for c in range(ncols):
                outs[c][rdisp[src]:rdisp[src] + len(sel)] = columns_per_rank[src][c][sel]
        received.append(outs)
    return received, send_counts


def shuffle_strings(strings_per_rank, dests_per_rank):
    """String column through the shuffle in the reference's wire format (shuffle_utils.py:277-297): every rank sends a
    LENGTH array and the concatenated characters per destination; the receiver rebuilds the offsets with a prefix sum
    (convert_len_arr_to_offset, sdc/native/str_ext/sdc_str_arr.hpp:242-252).
    strings_per_rank[r] = list of bytes objects -> received[r] = (offsets uint32[m + 1], chars uint8[total])."""
    P = len(strings_per_rank)
    out = []
    for r in range(P):
        lens, chunks = [], []
        for src in range(P):
            for s, d in zip(strings_per_rank[src], dests_per_rank[src]):
                if d == r:
                    lens.append(len(s))
                    chunks.append(s)
        offsets = np.zeros(len(lens) + 1, dtype=np.uint32)
        offsets[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64)) if lens else []
        out.append((offsets, np.frombuffer(b"".join(chunks), dtype=np.uint8)))
    return out
