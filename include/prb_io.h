/* prb_io.h -- C ABI of the file boundary of the B200-native Poisson reconstruction.
 *
 * The reference's only external contract is "files in, file out" (SURVEY.md 8b):
 *
 *   reference (file:line)                                             entry point here
 *   ----------------------------------------------------------------  ----------------------
 *   input dispatch by extension                main.cu:517-520        prbio_read_points
 *     ASCIIPointStream  "x y z nx ny nz" lines PointStream.inl:29-52
 *     BinaryPointStream raw float32 x6 .bnpts  PointStream.inl:53-92
 *     PLYPointStream    vertex x,y,z,nx,ny,nz  PointStream.inl:160-239, plyfile.cu:782-1039
 *   PlyWriteTriangles(out,&mesh,PLY_ASCII,center,scale)
 *                                              main.cu:4566, plyfile.cu:2769-2837,
 *                                              "%g " items plyfile.cu:2136-2141
 *                                                                     prbio_write_mesh
 *
 * Host-only code (no CUDA call is made by these functions); plain pointers and sizes.
 * Error behaviour: the reference prints a message and exit(0)s; here the same message text is
 * returned through prb_last_error() together with a negative status so that the CLI can print
 * it and exit, and a library caller can recover.
 */
#ifndef PRB_IO_H_
#define PRB_IO_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Reads oriented points.  Format by extension like main.cu:517-520: ".ply" (ascii,
 * binary_little_endian, binary_big_endian; properties matched by NAME x y z nx ny nz, any
 * scalar type, any order, extra properties skipped; `vertex` must be the first element),
 * ".bnpts" (raw float32 x6), anything else ASCII "x y z nx ny nz" records.  On success *xyz and
 * *normals are malloc'ed float32 [n][3] arrays owned by the caller (prbio_free). */
int prbio_read_points(const char* path, float** xyz, float** normals, int64_t* n);
void prbio_free(void* p);

/* Writes a triangle mesh as PLY: element vertex (float x y z) + element face
 * (list uchar int vertex_indices), vertices transformed as v*scale + center[] in float like
 * plyfile.cu:2801-2803.  binary = 0: ASCII, byte-identical to the reference writer ("%g " per
 * item, one element per line); binary = 1: binary_little_endian (fast path, not offered by
 * the reference's main()).  ".ply" is appended to a path that lacks it (plyfile.cu:251-257). */
int prbio_write_mesh(const char* path, const float* vertices, int64_t nv, const int32_t* triangles, int64_t nt,
                     const float center[3], float scale, int binary);

/* Optional (the reference never does this; `poisson_recon --weld`): merges vertices with bit-identical positions -- the seam
 * vertices every pass duplicates, insertTriangle main.cu:3220-3245 -- keeping first occurrences in order, and re-indexes the
 * triangles in place.  *nv_out = number of vertices left. */
int prbio_weld_mesh(float* vertices, int64_t nv, int32_t* triangles, int64_t nt, int64_t* nv_out);

#ifdef __cplusplus
}
#endif
#endif /* PRB_IO_H_ */
