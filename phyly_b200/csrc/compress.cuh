/*
 * Site-pattern compression on the device: identical alignment columns (rows of the [site][node] code array) are
 * merged into one pattern with a multiplicity, the input of every site-weighted query (site_reduction with weights,
 * reduction.c:24-118).  The reference leaves this to its input generators, which build an insertion-ordered dictionary
 * column -> count (examples/BEAST.GTRG/mknuc.py:57-66); this restates exactly that: patterns in order of first
 * occurrence, integer counts, plus the site -> pattern map needed to scatter per-pattern results back to sites.
 * Integer work only -- bit-exact against the oracle by construction.
 *
 *   1. pc_insert_kernel   one thread per site: 64-bit FNV-1a hash of the row, insertion into an open-addressing table
 *                         (linear probing, CAS on the slot's owner site, FULL row comparison on every hit: no false merges)
 *   2. pc_first_kernel    the owner of a slot becomes the smallest site index that landed there (atomicMin)
 *   3. scan               exclusive prefix sum of the "first occurrence" flags = pattern numbers in site order
 *   4. pc_emit_kernel     site -> pattern map, counts (integer atomics), rows of the first occurrences
 */
#pragma once
#include <stdint.h>

#define PC_EMPTY (-1)

__device__ __forceinline__ unsigned long long pc_hash_row(const unsigned char *row, int nbytes)
{
    unsigned long long h = 1469598103934665603ULL;
    for (int k = 0; k < nbytes; k++) { h ^= row[k]; h *= 1099511628211ULL; }
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
    return h;
}

__device__ __forceinline__ bool pc_rows_equal(const unsigned char *a, const unsigned char *b, int nbytes)
{
    for (int k = 0; k < nbytes; k++) if (a[k] != b[k]) return false;
    return true;
}

__global__ void pc_insert_kernel(const unsigned char *codes, int64_t S, int row_bytes, int *table, unsigned long long mask, int *slot_of)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const unsigned char *row = codes + (size_t)s * row_bytes;
    unsigned long long pos = pc_hash_row(row, row_bytes) & mask;
    while (true) {
        int owner = table[pos];
        if (owner == PC_EMPTY) {
            owner = atomicCAS(&table[pos], PC_EMPTY, (int)s);
            if (owner == PC_EMPTY) break;                      /* this site opened the slot */
        }
        if (pc_rows_equal(row, codes + (size_t)owner * row_bytes, row_bytes)) break;
        pos = (pos + 1) & mask;
    }
    slot_of[s] = (int)pos;
}

__global__ void pc_first_kernel(int64_t S, const int *slot_of, int *first_site)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) atomicMin(&first_site[slot_of[s]], (int)s);
}

/* flags[s] = 1 when site s is the first occurrence of its row */
__global__ void pc_flag_kernel(int64_t S, const int *slot_of, const int *first_site, int *flags)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) flags[s] = (first_site[slot_of[s]] == (int)s) ? 1 : 0;
}

/* exclusive scan in three steps: per-block sums (1024 items per block), scan of the sums by one block, offsets */
__global__ void pc_block_sum_kernel(const int *in, int64_t n, int *block_sums)
{
    __shared__ int red[32];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    int v = (i < n) ? in[i] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
    }
}

__global__ void pc_scan_sums_kernel(int *block_sums, int nblocks, int *total)
{
    /* one block, sequential over chunks of 1024: nblocks is S / 1024, small */
    __shared__ int buf[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < nblocks) ? block_sums[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblocks) block_sums[i] = carry + buf[threadIdx.x] - v;      /* exclusive */
        __syncthreads();
        if (threadIdx.x == 1023) carry += buf[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void pc_scan_final_kernel(const int *flags, int64_t n, const int *block_offsets, int *out)
{
    __shared__ int buf[1024];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int v = (i < n) ? flags[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
        __syncthreads();
        buf[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) out[i] = block_offsets[blockIdx.x] + buf[threadIdx.x] - v;
}

__global__ void pc_emit_kernel(const unsigned char *codes, int64_t S, int row_bytes, const int *slot_of, const int *first_site,
                               const int *pattern_of_site_scan, const int *flags, int *site_to_pattern, int *counts,
                               unsigned char *codes_out)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int rep = first_site[slot_of[s]];
    const int pid = pattern_of_site_scan[rep];
    site_to_pattern[s] = pid;
    atomicAdd(&counts[pid], 1);
    if (flags[s]) {
        const unsigned char *row = codes + (size_t)s * row_bytes;
        unsigned char *dst = codes_out + (size_t)pid * row_bytes;
        for (int k = 0; k < row_bytes; k++) dst[k] = row[k];
    }
}
