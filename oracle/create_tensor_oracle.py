"""CPU restatement of the reference's CreateTensor.  TEST INFRASTRUCTURE - only tests/, smoke() and
bench.py's cpu_baseline leg may import this; the product path is clair_b200/create_tensor.py over the CUDA kernel.

Follows /root/reference/dataPrepScripts/CreateTensor.py:
    OutputAlnTensor   :179-394   SAM rows -> per-centre alignment records (MQ filter :264, depth cap :274-281,
                                 CIGAR walk :283-366, active-set bookkeeping :298-320,350-360)
    generate_tensor   :29-65     records -> 33x8x4 counts, coverage / left-edge guard :54-56, row text :57-62
    candidate_position_generator_from :68-105  (windows a centre is opened by: :88-96)
Differences in form, not in result: records are folded into the counts as they are produced instead of being stored per
read, and the rows come back as tuples (the text row is `format_row`).  The 5,000,000-record memory guard
(`available_slots`, :180,285-286,306-308) is NOT modelled: it only drops records when a region holds more than five million
outstanding records.  Parity is PINNED: tests/golden/create_tensor_cases.json.gz holds rows printed by the reference's own
OutputAlnTensor (oracle/gen_golden_create_tensor.py) and this file reproduces every one of them.
"""
import numpy as np

FLANK = 16                      # shared/param.py:9
N_POS = 2 * FLANK + 1           # CreateTensor.py:24
# shared/utils.py:24-27
BASE2NUM = dict(zip("ACGTURYSWKMBDHVN", (0, 1, 2, 3, 3, 0, 1, 1, 0, 2, 0, 1, 0, 0, 0, 0)))


def _add_record(counts, depth, center, ref_pos, query_adv, ref_base, query_base, strand):
    """One alignment record folded into a centre's counts: generate_tensor's loop body (CreateTensor.py:36-52)."""
    if (ref_base != "-" and ref_base not in BASE2NUM) or (query_base != "-" and query_base not in BASE2NUM):
        return
    idx = ref_pos - center + (FLANK + 1)
    if not 0 <= idx < N_POS:
        return
    off = 4 if strand else 0
    if query_base != "-" and ref_base != "-":
        depth[idx] += 1
        counts[idx, BASE2NUM[ref_base] + off, 0] += 1
        counts[idx, BASE2NUM[query_base] + off, 1] += 1
        counts[idx, BASE2NUM[ref_base] + off, 2] += 1
        counts[idx, BASE2NUM[query_base] + off, 3] += 1
    elif query_base != "-":
        counts[min(idx + query_adv, N_POS - 1), BASE2NUM[query_base] + off, 1] += 1
    else:
        counts[idx, BASE2NUM[ref_base] + off, 2] += 1


def create_tensors(sam_lines, candidate_positions, reference_sequence, reference_start_0_based=0, ctg_name="chr",
                   min_mq=0, dcov=250, min_coverage=0, consider_left_edge=True, ctg_start=None, ctg_end=None):
    """-> list of (ctg_name, centre (1-based), 33-base reference window, counts int32 [33,8,4]) in output order.

    sam_lines: rows of `samtools view` (header rows allowed); candidate_positions: 1-based, ascending;
    reference_sequence: the upper-cased text `samtools faidx` returned, starting at reference_start_0_based."""
    opens = {}                      # 0-based reference position -> centres a read opens when it aligns a base there
    for position in candidate_positions:
        position = int(position)
        if ctg_start is not None and ctg_end is not None and not ctg_start <= position <= ctg_end:
            continue
        if consider_left_edge:
            for i in range(position - (FLANK + 1), position + (FLANK + 1)):
                opens.setdefault(i, []).append(position)
        else:
            opens[position - (FLANK + 1)] = [position]

    tensors = {}                    # centre -> (counts, depth); insertion order is the output order
    previous_position, depth_cap = 0, 0
    for line in sam_lines:
        col = line.split()
        if col[0][0] == "@":
            continue
        flag, pos, mq, cigar, seq = int(col[1]), int(col[3]) - 1, int(col[4]), col[5], col[9].upper()
        strand = (flag & 16) == 16
        if mq < min_mq:
            continue
        if previous_position != pos:
            previous_position, depth_cap = pos, 0
        else:
            depth_cap += 1
            if depth_cap >= dcov:
                continue

        active = set()

        def open_at(p):
            for c in opens.get(p, ()):
                if c not in active:
                    active.add(c)
                    tensors.setdefault(c, (np.zeros((N_POS, 8, 4), np.int32), [0] * N_POS))

        def close_at(p):
            active.discard(p - (FLANK + 1))

        ref_pos, q_pos, n = pos, 0, 0
        for ch in cigar:
            if ch.isdigit():
                n = n * 10 + int(ch)
                continue
            if ch == "S":
                q_pos += n
            elif ch in "M=X":
                for _ in range(n):
                    open_at(ref_pos)
                    rb = reference_sequence[ref_pos - reference_start_0_based] if active else None
                    for c in active:
                        _add_record(tensors[c][0], tensors[c][1], c, ref_pos, 0, rb, seq[q_pos], strand)
                    close_at(ref_pos)
                    ref_pos += 1
                    q_pos += 1
            elif ch == "I":
                for adv in range(n):
                    for c in active:
                        _add_record(tensors[c][0], tensors[c][1], c, ref_pos, adv, "-", seq[q_pos], strand)
                    q_pos += 1
            elif ch == "D":
                for _ in range(n):
                    rb = reference_sequence[ref_pos - reference_start_0_based] if active else None
                    for c in active:
                        _add_record(tensors[c][0], tensors[c][1], c, ref_pos, 0, rb, "-", strand)
                    open_at(ref_pos)
                    close_at(ref_pos)
                    ref_pos += 1
            n = 0

    rows = []
    for center, (counts, depth) in tensors.items():
        new_ref = center - reference_start_0_based
        if new_ref - (FLANK + 1) < 0 or depth[FLANK] < min_coverage:
            continue
        rows.append((ctg_name, center, reference_sequence[new_ref - (FLANK + 1):new_ref + FLANK], counts))
    return rows


def format_row(row):
    """The text row the reference prints for one tensor (CreateTensor.py:57-62)."""
    ctg, center, seq, counts = row
    return "%s %d %s %s" % (ctg, center, seq, " ".join("%d" % v for v in counts.reshape(-1)))
