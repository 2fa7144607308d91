// shapes.cuh -- compile-time row-run structuring elements shared by the
// disk (k_disk.cu) and median (k_median.cu) kernels.  A shape is its reach R
// and the half-width hw(dy+R) of the centred run on each row; the dispatcher
// matches the run-length form of the caller's element list against this table.
#pragma once

template <int ID> struct Shape;
#define MORSI_SHAPE(ID, RY, ...) \
	template <> struct Shape<ID> { \
		static constexpr int R = RY; \
		__host__ __device__ static constexpr int hw(int i) { constexpr int t[2 * RY + 1] = {__VA_ARGS__}; return t[i]; } \
	};
MORSI_SHAPE(0, 2, 1, 2, 2, 2, 1)                                    // disk2.5
MORSI_SHAPE(1, 2, 2, 2, 2, 2, 2)                                    // disk3 (5x5)
MORSI_SHAPE(2, 3, 1, 2, 3, 3, 3, 2, 1)                              // disk3.5
MORSI_SHAPE(3, 3, 2, 3, 3, 3, 3, 3, 2)                              // disk4
MORSI_SHAPE(4, 4, 1, 2, 3, 4, 4, 4, 3, 2, 1)                        // disk4.2
MORSI_SHAPE(5, 4, 2, 3, 4, 4, 4, 4, 4, 3, 2)                        // disk4.5, disk5
MORSI_SHAPE(6, 5, 1, 3, 4, 4, 5, 5, 5, 4, 4, 3, 1)                  // disk5.1
MORSI_SHAPE(7, 5, 3, 4, 5, 5, 5, 5, 5, 5, 5, 4, 3)                  // disk6
MORSI_SHAPE(8, 6, 3, 4, 5, 6, 6, 6, 6, 6, 6, 6, 5, 4, 3)            // disk7
MORSI_SHAPE(9, 7, 3, 5, 6, 6, 7, 7, 7, 7, 7, 7, 7, 6, 6, 5, 3)      // disk8
MORSI_SHAPE(10, 8, 4, 5, 6, 7, 8, 8, 8, 8, 8, 8, 8, 8, 8, 7, 6, 5, 4)   // disk9
MORSI_SHAPE(11, 9, 4, 5, 7, 7, 8, 9, 9, 9, 9, 9, 9, 9, 9, 9, 8, 7, 7, 5, 4)   // disk10
MORSI_SHAPE(12, 11, 4, 6, 7, 8, 9, 10, 10, 11, 11, 11, 11, 11, 11, 11, 11, 11, 10, 10, 9, 8, 7, 6, 4)   // disk12
MORSI_SHAPE(13, 14, 5, 7, 8, 10, 11, 11, 12, 13, 13, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 13, 13, 12, 11, 11, 10, 8, 7, 5)   // disk15
#define MORSI_NSHAPES 14

MORSI_SHAPE(14, 1, 0, 1, 0)                                         // cross (as a set; element order differs)
MORSI_SHAPE(15, 1, 1, 1, 1)                                         // square, disk2
// more disks (k_disk only)
MORSI_SHAPE(16, 10, 4, 6, 7, 8, 9, 9, 10, 10, 10, 10, 10, 10, 10, 10, 10, 9, 9, 8, 7, 6, 4)   // disk11
MORSI_SHAPE(17, 12, 4, 6, 8, 9, 10, 10, 11, 11, 12, 12, 12, 12, 12, 12, 12, 12, 12, 11, 11, 10, 10, 9, 8, 6, 4)   // disk13
MORSI_SHAPE(18, 13, 5, 7, 8, 9, 10, 11, 12, 12, 13, 13, 13, 13, 13, 13, 13, 13, 13, 13, 13, 12, 12, 11, 10, 9, 8, 7, 5)   // disk14
