// TMEM <-> register transfers in the 16-lane shapes (generated operand lists; layouts from the PTX ISA, cross-checked
// against the (thread, value) -> bit maps of the vendor's copy traits):
//   16x256b.xN  thread t of the warp, register 4n + 2h + e  <->  TMEM lane (base + t/4 + 8h), column 8n + 2(t%4) + e
//   16x128b.xN  thread t, register 2n + h                   <->  TMEM lane (base + t/4 + 8h), column 4n + (t%4)
// i.e. the m16n8 accumulator fragment: a row's columns are spread over the 4 threads of a quad, so a warp that owns 16
// rows has whole rows to itself and row reductions are two quad shuffles. fp32 pairs (columns 8n + 2q, +1) pack into
// the bf16x2 word of column 4n + q: the 16x128b store of the packed P tile needs no data movement between threads.
#pragma once
#include <stdint.h>

#define MAGI_TMEM_LD_16x256b_X16(taddr, r, o) \
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];" \
    : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15]), "=r"(r[o + 16]), "=r"(r[o + 17]), "=r"(r[o + 18]), "=r"(r[o + 19]), "=r"(r[o + 20]), "=r"(r[o + 21]), "=r"(r[o + 22]), "=r"(r[o + 23]), "=r"(r[o + 24]), "=r"(r[o + 25]), "=r"(r[o + 26]), "=r"(r[o + 27]), "=r"(r[o + 28]), "=r"(r[o + 29]), "=r"(r[o + 30]), "=r"(r[o + 31]), "=r"(r[o + 32]), "=r"(r[o + 33]), "=r"(r[o + 34]), "=r"(r[o + 35]), "=r"(r[o + 36]), "=r"(r[o + 37]), "=r"(r[o + 38]), "=r"(r[o + 39]), "=r"(r[o + 40]), "=r"(r[o + 41]), "=r"(r[o + 42]), "=r"(r[o + 43]), "=r"(r[o + 44]), "=r"(r[o + 45]), "=r"(r[o + 46]), "=r"(r[o + 47]), "=r"(r[o + 48]), "=r"(r[o + 49]), "=r"(r[o + 50]), "=r"(r[o + 51]), "=r"(r[o + 52]), "=r"(r[o + 53]), "=r"(r[o + 54]), "=r"(r[o + 55]), "=r"(r[o + 56]), "=r"(r[o + 57]), "=r"(r[o + 58]), "=r"(r[o + 59]), "=r"(r[o + 60]), "=r"(r[o + 61]), "=r"(r[o + 62]), "=r"(r[o + 63]) \
    : "r"(taddr) : "memory")

#define MAGI_TMEM_ST_16x256b_X16(taddr, r, o) \
  asm volatile("tcgen05.st.sync.aligned.16x256b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};" \
    :: "r"(taddr), "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7]), "r"(r[o + 8]), "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]), "r"(r[o + 12]), "r"(r[o + 13]), "r"(r[o + 14]), "r"(r[o + 15]), "r"(r[o + 16]), "r"(r[o + 17]), "r"(r[o + 18]), "r"(r[o + 19]), "r"(r[o + 20]), "r"(r[o + 21]), "r"(r[o + 22]), "r"(r[o + 23]), "r"(r[o + 24]), "r"(r[o + 25]), "r"(r[o + 26]), "r"(r[o + 27]), "r"(r[o + 28]), "r"(r[o + 29]), "r"(r[o + 30]), "r"(r[o + 31]), "r"(r[o + 32]), "r"(r[o + 33]), "r"(r[o + 34]), "r"(r[o + 35]), "r"(r[o + 36]), "r"(r[o + 37]), "r"(r[o + 38]), "r"(r[o + 39]), "r"(r[o + 40]), "r"(r[o + 41]), "r"(r[o + 42]), "r"(r[o + 43]), "r"(r[o + 44]), "r"(r[o + 45]), "r"(r[o + 46]), "r"(r[o + 47]), "r"(r[o + 48]), "r"(r[o + 49]), "r"(r[o + 50]), "r"(r[o + 51]), "r"(r[o + 52]), "r"(r[o + 53]), "r"(r[o + 54]), "r"(r[o + 55]), "r"(r[o + 56]), "r"(r[o + 57]), "r"(r[o + 58]), "r"(r[o + 59]), "r"(r[o + 60]), "r"(r[o + 61]), "r"(r[o + 62]), "r"(r[o + 63]) : "memory")

#define MAGI_TMEM_LD_16x256b_X8(taddr, r, o) \
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
    : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15]), "=r"(r[o + 16]), "=r"(r[o + 17]), "=r"(r[o + 18]), "=r"(r[o + 19]), "=r"(r[o + 20]), "=r"(r[o + 21]), "=r"(r[o + 22]), "=r"(r[o + 23]), "=r"(r[o + 24]), "=r"(r[o + 25]), "=r"(r[o + 26]), "=r"(r[o + 27]), "=r"(r[o + 28]), "=r"(r[o + 29]), "=r"(r[o + 30]), "=r"(r[o + 31]) \
    : "r"(taddr) : "memory")

#define MAGI_TMEM_ST_16x256b_X8(taddr, r, o) \
  asm volatile("tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" \
    :: "r"(taddr), "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7]), "r"(r[o + 8]), "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]), "r"(r[o + 12]), "r"(r[o + 13]), "r"(r[o + 14]), "r"(r[o + 15]), "r"(r[o + 16]), "r"(r[o + 17]), "r"(r[o + 18]), "r"(r[o + 19]), "r"(r[o + 20]), "r"(r[o + 21]), "r"(r[o + 22]), "r"(r[o + 23]), "r"(r[o + 24]), "r"(r[o + 25]), "r"(r[o + 26]), "r"(r[o + 27]), "r"(r[o + 28]), "r"(r[o + 29]), "r"(r[o + 30]), "r"(r[o + 31]) : "memory")

#define MAGI_TMEM_LD_16x256b_X4(taddr, r, o) \
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
    : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15]) \
    : "r"(taddr) : "memory")

#define MAGI_TMEM_ST_16x256b_X4(taddr, r, o) \
  asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" \
    :: "r"(taddr), "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7]), "r"(r[o + 8]), "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]), "r"(r[o + 12]), "r"(r[o + 13]), "r"(r[o + 14]), "r"(r[o + 15]) : "memory")

#define MAGI_TMEM_ST_16x128b_X16(taddr, r, o) \
  asm volatile("tcgen05.st.sync.aligned.16x128b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" \
    :: "r"(taddr), "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7]), "r"(r[o + 8]), "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]), "r"(r[o + 12]), "r"(r[o + 13]), "r"(r[o + 14]), "r"(r[o + 15]), "r"(r[o + 16]), "r"(r[o + 17]), "r"(r[o + 18]), "r"(r[o + 19]), "r"(r[o + 20]), "r"(r[o + 21]), "r"(r[o + 22]), "r"(r[o + 23]), "r"(r[o + 24]), "r"(r[o + 25]), "r"(r[o + 26]), "r"(r[o + 27]), "r"(r[o + 28]), "r"(r[o + 29]), "r"(r[o + 30]), "r"(r[o + 31]) : "memory")

#define MAGI_TMEM_LD_16x128b_X16(taddr, r, o) \
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
    : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15]), "=r"(r[o + 16]), "=r"(r[o + 17]), "=r"(r[o + 18]), "=r"(r[o + 19]), "=r"(r[o + 20]), "=r"(r[o + 21]), "=r"(r[o + 22]), "=r"(r[o + 23]), "=r"(r[o + 24]), "=r"(r[o + 25]), "=r"(r[o + 26]), "=r"(r[o + 27]), "=r"(r[o + 28]), "=r"(r[o + 29]), "=r"(r[o + 30]), "=r"(r[o + 31]) \
    : "r"(taddr) : "memory")

#define MAGI_TMEM_ST_16x128b_X8(taddr, r, o) \
  asm volatile("tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" \
    :: "r"(taddr), "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7]), "r"(r[o + 8]), "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]), "r"(r[o + 12]), "r"(r[o + 13]), "r"(r[o + 14]), "r"(r[o + 15]) : "memory")

#define MAGI_TMEM_LD_16x128b_X8(taddr, r, o) \
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
    : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15]) \
    : "r"(taddr) : "memory")

#define MAGI_TMEM_LD_16x256b_X2(taddr, r, o) \
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" \
    : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]) \
    : "r"(taddr) : "memory")

#define MAGI_TMEM_ST_16x256b_X2(taddr, r, o) \
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" \
    :: "r"(taddr), "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7]) : "memory")
