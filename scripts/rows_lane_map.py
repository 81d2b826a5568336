"""Host-side check of the thread -> (row, channel quad) mapping of the gemm_rows producers (csrc/mlp_tc.cu): every
(row, quad) of a 128 x 32 chunk is staged exactly once, and a warp-wide 128-bit tile store touches each 16-byte bank
group once per quarter-warp (the first mapping, k4 = tid & 7, put all 8 lanes of a quarter-warp on the same group:
ncu l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st = 37 M of 45 M store wavefronts, profiles/r01_ncu_source_gemm.md)."""
KC = 32


def st_off(r, k4):                       # byte offset inside the canonical K-major tile (8 x 16 B core matrices)
    return ((r >> 3) * (KC // 4) * 32 + k4 * 32 + (r & 7) * 4) * 4


def mapping(tid):                        # the kernel's expression
    return (tid >> 3) & 7, (tid >> 6) * 8 + (tid & 7)


def old_mapping(tid):
    return tid & 7, tid >> 3


def check(fn, threads, rows_per_thread, step):
    seen, worst = set(), 0
    for tid in range(threads):
        k4, rsub = fn(tid)
        for j in range(rows_per_thread):
            seen.add((rsub + step * j, k4))
    for w in range(threads // 32):
        for j in range(rows_per_thread):
            for q in range(4):
                groups = {}
                for lane in range(q * 8, q * 8 + 8):
                    k4, rsub = fn(w * 32 + lane)
                    g = (st_off(rsub + step * j, k4) // 16) % 8
                    groups[g] = groups.get(g, 0) + 1
                worst = max(worst, max(groups.values()))
    return len(seen), worst


if __name__ == "__main__":
    print("async path, new mapping: covered", *check(mapping, 256, 4, 32), "(items, worst lanes per bank group)")
    print("sync path,  new mapping: covered", *check(mapping, 128, 8, 16))
    print("async path, old mapping: covered", *check(old_mapping, 256, 4, 32))
