"""oracle/dpu_path.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement of the reference's *PIM host orchestration* (sort-merge-join/app.c) with each DPU program replaced
by its DPU-level contract, so that the stage-by-stage intermediates of the DPU path -- which rows every DPU selects,
which chunk it sorts, what every merge round pairs, how the join ranges are cut -- can be compared with the
single-pass CPU path (cpu_app.c) and with the B200 stages (smj_select / smj_sort / smj_merge / smj_join).

Why it exists: BASELINE.json asks for a cross-check of the DPU path on small inputs under the UPMEM functional
simulator; the simulator library (libdpufsim) is missing from the reference checkout (.MISSING_LARGE_BLOBS:8) and the
DPU toolchain cannot run here, so the orchestration is restated instead (SURVEY.md section 8f
item 4).  PARITY UNPINNED against a running DPU system -- there is none to run -- but pinned where it can be: for
unique join keys the reference's design makes the DPU path's result.csv equal cpu_app.c's (per-DPU joins over disjoint
ascending key ranges written in DPU order, app.c:739-753), and tests/test_dpu_path.py checks exactly that against the
golden vectors cpu_app.c produced (tests/golden/).

What is restated, with the reference lines it follows:
  partition   app.c:155-218   row blocks over NR_DPUS: table 0 first, `row_size = (r1 + r2) / NR_DPUS` rows per DPU
  select      app.c:221-288   per-DPU filter (select.c:24-39, order kept), then concatenation per table in DPU order
  sort        app.c:315-373   even re-split of the selected rows (last DPU takes the remainder), per-DPU sort
                              (sort_dpu.c:157-187 insertion sort per tasklet + :251-323 tasklet merge tree)
  merge       app.c:413-547   log2 tournament: runs (0,1), (2,3), ... merged by merge_dpu.c, an odd last run carried
  join split  app.c:585-633   table 1 cut into pivot_id even chunks; table 2 cut where app.c:94-121 `binary_search`
                              finds each chunk's last key (ANY equal index, or the last smaller one)
  join        app.c:638-688   per-DPU zipper join of (chunk i, slice i) (join.c:58-266), results in DPU order

DPU-level contracts assumed for the device programs (the reports only ever ran unique keys):
  * sort_dpu / merge_dpu order equal keys stably (left run first).  The DPU code's tasklet split by `binary_search`
    can cut a run of equal keys anywhere, so with duplicate keys the real DPU order is not defined; neither is it here.
  * select.c:73-74 reads SELECT_VAL into an `unsigned int`: a negative knob becomes val + 2^32 and nothing passes.
    `unsigned_select_val=True` (default) reproduces that, False gives cpu_app.c's signed compare.
Reference defects that are NOT reproduced (they read outside the arrays): the last table-2 slice is sized
`total_row_num2 - cur_idx_t2 + 1` (app.c:628-629), one row past the end of the merged table; the emulator uses the
rows that exist.  With fewer selected table-1 rows than DPUs app.c:603 reads the key of row -1; the emulator gives
such empty chunks an empty table-2 slice.  `pivot_id` stays -1 when table 1 alone fills every DPU but the last
(app.c:186-199); the emulator raises ValueError for such shapes.
"""
import numpy as np


def _binary_search(keys, target):
    """app.c:94-121: index of ANY row whose key equals target, else the last index with key < target, else -1."""
    left, right, idx = 0, len(keys) - 1, -1
    while left <= right:
        mid = (left + right) // 2
        k = keys[mid]
        if k == target:
            return mid
        if k < target:
            idx = mid
            left = mid + 1
        else:
            right = mid - 1
    return idx


def _zipper(l, r, key1, key2):
    """join.c / cpu_app.c:204-266 at DPU level: both cursors advance on equal keys; row = T1 cols, T2 cols but key2."""
    keep = [c for c in range(r.shape[1]) if c != key2]
    i = j = 0
    li, rj = [], []
    kl, kr = l[:, key1], r[:, key2]
    while i < len(kl) and j < len(kr):
        if kl[i] == kr[j]:
            li.append(i); rj.append(j)
            i += 1; j += 1
        elif kl[i] < kr[j]:
            i += 1
        else:
            j += 1
    if not li:
        return np.empty((0, l.shape[1] + len(keep)), np.int32)
    return np.concatenate([l[li], r[rj][:, keep]], axis=1).astype(np.int32)


def _stable_sort(t, key):
    return t[np.argsort(t[:, key], kind="stable")]


def _merge(a, b, key):
    """merge_dpu.c at DPU level: two sorted runs of the same table into one, run a first on equal keys."""
    both = np.concatenate([a, b], axis=0)
    return both[np.argsort(both[:, key], kind="stable")]


def partition(r1, r2, nr_dpus):
    """app.c:155-218.  Returns (blocks, pivot_id, row_size): blocks[d] = (table_num, first_row, rows)."""
    row_size = (r1 + r2) // nr_dpus
    if row_size == 0:
        using, row_size = 2, r1
    else:
        using = nr_dpus
    blocks, pivot_id = [], -1
    first, second = r1, r2
    for i in range(using - 1):
        if first > 0:
            if first > row_size:
                rows = row_size
                first -= row_size
            else:
                rows = first
                first = 0
                pivot_id = i + 1
            blocks.append((0, i * row_size, rows))
        else:
            rows = row_size if second >= row_size else second
            second -= rows
            blocks.append((1, (i - pivot_id) * row_size, rows))
    if pivot_id < 1:
        raise ValueError("app.c:186-199 leaves pivot_id unset for this shape (table 1 does not end before the last DPU)")
    blocks.append((1, (using - 1 - pivot_id) * row_size, second))
    return blocks, pivot_id, row_size


def run(t1, t2, sel_col1=0, sel_val1=5000, sel_col2=0, sel_val2=5000, key1=0, key2=0, nr_dpus=64,
        unsigned_select_val=True):
    """The whole DPU path.  Returns a dict of stage intermediates; ['result'] is what app.c:720-755 writes."""
    t1 = np.ascontiguousarray(t1, dtype=np.int32)
    t2 = np.ascontiguousarray(t2, dtype=np.int32)
    tabs = (t1, t2)
    sel_col, sel_val, key = (sel_col1, sel_col2), [int(sel_val1), int(sel_val2)], (key1, key2)
    if unsigned_select_val:
        sel_val = [v & 0xFFFFFFFF for v in sel_val]          # select.c:73-74
    out = {}

    # ---- partition + select (app.c:155-288)
    blocks, pivot_id, _ = partition(t1.shape[0], t2.shape[0], nr_dpus)
    using = len(blocks)
    out["blocks"], out["pivot_id"] = blocks, pivot_id
    sel = []
    for tn, row0, rows in blocks:
        blk = tabs[tn][row0:row0 + rows]
        sel.append(blk[blk[:, sel_col[tn]].astype(np.int64) > sel_val[tn]])
    out["select_per_dpu"] = sel
    selected = [np.concatenate(sel[:pivot_id], axis=0), np.concatenate(sel[pivot_id:], axis=0)]
    out["selected"] = selected
    total = [selected[0].shape[0], selected[1].shape[0]]

    # ---- sort (app.c:315-373): even re-split, last DPU of each table takes the remainder
    ndpu = [pivot_id, using - pivot_id]
    chunks = [[], []]
    for tn in (0, 1):
        rs = total[tn] // ndpu[tn]
        for d in range(ndpu[tn]):
            lo = d * rs
            hi = (d + 1) * rs if d < ndpu[tn] - 1 else total[tn]
            chunks[tn].append(_stable_sort(selected[tn][lo:hi], key[tn]))
    out["sorted_chunks"] = [list(c) for c in chunks]

    # ---- merge tournament (app.c:413-547)
    rounds = []
    runs = [list(chunks[0]), list(chunks[1])]
    done = [False, False]
    while not (done[0] and done[1]):
        this_round = []
        for tn in (0, 1):
            if done[tn]:
                continue
            cur = runs[tn]
            nxt = [_merge(cur[p], cur[p + 1], key[tn]) for p in range(0, len(cur) - 1, 2)]
            this_round.append((tn, [(cur[p].shape[0], cur[p + 1].shape[0]) for p in range(0, len(cur) - 1, 2)]))
            if len(cur) % 2 == 1:
                nxt.append(cur[-1])                           # app.c:505-520: the odd run is carried
            runs[tn] = nxt
        for tn in (0, 1):
            if len(runs[tn]) <= 1:
                done[tn] = True
        rounds.append(this_round)
    out["merge_rounds"] = rounds
    merged = [runs[0][0], runs[1][0]]
    out["merged"] = merged

    # ---- join range split (app.c:585-633)
    rs = total[0] // pivot_id
    k2 = merged[1][:, key[1]]
    l_chunks, r_slices, used_idx, cur = [], [], [], 0
    for i in range(pivot_id):
        lo = i * rs
        hi = (i + 1) * rs if i < pivot_id - 1 else total[0]
        l_chunks.append(merged[0][lo:hi])
        if i < pivot_id - 1:
            if rs == 0:
                # fewer selected table-1 rows than DPUs: app.c:603 reads the key of row -1 (before the array);
                # the emulator gives the empty chunk an empty slice, the last chunk then takes every table-2 row
                used_idx.append(cur - 1)
                r_slices.append(merged[1][cur:cur])
                continue
            u = _binary_search(k2, merged[0][hi - 1, key[0]])
            used_idx.append(u)
            r_slices.append(merged[1][cur:u + 1])
            cur = u + 1
        else:
            r_slices.append(merged[1][cur:])                  # (app.c:628 sizes this one row too long)
    out["join_chunks"], out["join_slices"], out["used_idx"] = l_chunks, r_slices, used_idx

    # ---- per-DPU join, results in DPU order (app.c:638-688, 739-753)
    joined = [_zipper(l_chunks[i], r_slices[i], key[0], key[1]) for i in range(pivot_id)]
    out["join_per_dpu"] = joined
    c_out = t1.shape[1] + t2.shape[1] - 1
    out["result"] = np.concatenate(joined, axis=0) if joined else np.empty((0, c_out), np.int32)
    return out
