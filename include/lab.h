/*
 * lab.h -- the lab's operator boundary, served by libb200sort.so.
 *
 * Same two C++-linkage prototypes as the reference's SRM/include/lab.h:9-10 (mangled
 * _Z11order_arrayPii / _Z16order_with_trustPii), so SRM/main.cpp and SRM/performanceTest.cpp link
 * against this library unchanged.  Contract:
 *   - `src` is a HOST array of `length` 32-bit signed keys, sorted ascending in place;
 *   - the call blocks until the result is in `src`;
 *   - the caller owns `src`; the library owns every device resource;
 *   - on any failure the call prints "GPUassert: <message> <file> <line>" to stderr and exits with
 *     a non-zero code, the reference's convention (SRM/include/utils.h:18-26).
 *
 * order_array       runs the onesweep LSD radix sort  (B200SORT_ALGO_RADIX).
 * order_with_trust  in the reference is a Thrust call the drivers time beside order_array
 *                   (SRM/lab.cu:404-406).  This library contains no Thrust/CUB and no host sort:
 *                   here it runs the merge sort (B200SORT_ALGO_MERGE) on the GPU, so the drivers'
 *                   two columns read "radix" and "merge".
 */
#ifndef B200SORT_LAB_H
#define B200SORT_LAB_H

#include "utils.h"

void order_array(int *srcCpu, int length);
void order_with_trust(int *src, int length);

#endif /* B200SORT_LAB_H */
