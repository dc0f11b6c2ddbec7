/* housescan_b200 — C ABI of the B200-native point-cloud path of nh2/housescan.
 *
 * This is the drop-in boundary.  The reference has no FFI for this path; the
 * boundary it does have is the export list of its Haskell helper modules.  Every entry point
 * below names the reference interface it replaces (file:line under /root/reference/housescan).
 * A Haskell maintainer binds these with `foreign import ccall safe` (see INTEGRATION.md and
 * haskell/*.hs); the tests and bench bind the identical symbols through Python ctypes.
 *
 * Conventions
 *   - plain C types only; all functions return int32_t status (HS_OK == 0) unless noted.
 *   - host buffers are caller-owned and never retained past return.
 *   - device clouds are opaque handles owned by the ctx, freed with hs_cloud_free.
 *   - one hs_ctx == one CUDA device + one stream; one in-flight call per ctx (not re-entrant);
 *     several ctxs (one per GPU, or one per process under torchrun) may coexist.
 *   - NO CPU FALLBACK: without an sm_100 device hs_ctx_create fails with HS_ECUDA.
 *   - points are AoS float xyz, 12 B/point, exactly `Vector Vec3` (Main.hs:39-42,120,641).
 *   - planes are 4 floats nx,ny,nz,d of `PlaneEq` n.x = d, |n| = 1 (Main.hs:1357).
 *   - 4x4 transforms are row-major, right-multiplied (p' = p .* M), translation in row 3
 *     (Main.hs:10, :1725-1730).
 */
#ifndef HOUSESCAN_B200_H
#define HOUSESCAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_API __attribute__((visibility("default")))

/* status codes; Haskell shim maps: HS_ECUDA/HS_EIO -> Left String (HoniHelper.hs:25-42),
 * HS_ESINGULAR -> Nothing (TranslationOptimizer.hs:66, Main.hs:2151), HS_EINVAL -> error "..." */
enum { HS_OK = 0, HS_EINVAL = 1, HS_ECUDA = 2, HS_ENCCL = 3, HS_ENOMEM = 4, HS_ESINGULAR = 5, HS_EIO = 6 };

typedef struct hs_ctx hs_ctx;
typedef struct hs_cloud hs_cloud;

#define HS_REC 24 /* doubles per room in a cuboid-sums record (layout in DESIGN.md) */
#define HS_PS 10  /* doubles per (room, plane) in hs_plane_sums */
#define HS_NE 29  /* doubles per frame in hs_backproject_reduce6x6: 21 JtJ + 6 Jtr + sum r^2 + count */

/* ---- context ------------------------------------------------------------------------------ */
HS_API int32_t hs_ctx_create(int32_t device, hs_ctx** out);
HS_API int32_t hs_ctx_destroy(hs_ctx* ctx);
HS_API const char* hs_last_error(const hs_ctx* ctx); /* ctx may be NULL: last error of a failed create */
HS_API int32_t hs_ctx_set_stream(hs_ctx* ctx, void* cuda_stream); /* borrow an external stream (NULL = own) */
HS_API int32_t hs_ctx_sync(hs_ctx* ctx);
HS_API int32_t hs_ctx_device(const hs_ctx* ctx);
HS_API int32_t hs_ctx_sm_count(const hs_ctx* ctx);
HS_API int64_t hs_ctx_launch_count(const hs_ctx* ctx); /* kernels this ctx has launched so far */
HS_API int32_t hs_ctx_set_mode(hs_ctx* ctx, int32_t key, int32_t value); /* tuning knobs, see DESIGN.md */

/* ---- clouds: `Cloud.cloudPoints :: Vector Vec3` (Main.hs:117-121) ------------------------- */
HS_API int32_t hs_cloud_upload(hs_ctx* ctx, const float* xyz_aos, int64_t n, hs_cloud** out);
HS_API int32_t hs_cloud_alloc(hs_ctx* ctx, int64_t n, hs_cloud** out);
HS_API int32_t hs_cloud_wrap_device(hs_ctx* ctx, void* device_xyz_aos, int64_t n, hs_cloud** out); /* borrowed, 16 B aligned */
HS_API int32_t hs_cloud_write(hs_ctx* ctx, hs_cloud* cloud, const float* xyz_aos, int64_t n);    /* async H2D into existing cloud */
HS_API int32_t hs_cloud_download(hs_ctx* ctx, const hs_cloud* cloud, float* xyz_out);
HS_API int64_t hs_cloud_size(const hs_cloud* cloud);
HS_API void* hs_cloud_device_ptr(const hs_cloud* cloud);
HS_API int32_t hs_cloud_free(hs_ctx* ctx, hs_cloud* cloud);

/* ---- (1) depth frames -> points.  HoniHelper.takeDepthSnapshot frame format (HoniHelper.hs:20,34-36,45-46)
 *      + Main.addDevicePointCloud (Main.hs:1296-1313): mask d != 0, order-preserving compaction,
 *      x/10, y/10, d/20 - 30.  xyz_out holds up to w*h points; mask_out w*h bytes (either may be NULL). */
HS_API int32_t hs_backproject_ref(hs_ctx* ctx, const uint16_t* depth, int32_t w, int32_t h, float* xyz_out,
                                  uint8_t* mask_out, int64_t* n_valid);
/* device-resident variant: frame already on the device, cloud_out preallocated with >= w*h points */
HS_API int32_t hs_backproject_ref_dev(hs_ctx* ctx, const void* d_depth, int32_t w, int32_t h, hs_cloud* cloud_out,
                                      void* d_mask_or_null, int64_t* n_valid);
/* fused per-frame back-projection + point-to-plane assignment + 6x6 normal equations (north-star piece (1)+(2);
 * no reference code).  intr = fx,fy,cx,cy or NULL (=> scalePoints of Main.hs:1311-1313); poses = nframes x 16 or NULL.
 * out = nframes x HS_NE doubles (host).  frames are host pointers here, device pointers in the _dev variant. */
HS_API int32_t hs_backproject_reduce6x6(hs_ctx* ctx, const uint16_t* frames, int64_t nframes, int32_t w, int32_t h,
                                        const float* intr, const float* poses, const float* planes, int32_t K, double* out);
HS_API int32_t hs_backproject_reduce6x6_dev(hs_ctx* ctx, const void* d_frames, int64_t nframes, int32_t w, int32_t h,
                                            const float* intr, const float* poses, const float* planes, int32_t K,
                                            void* d_out);

/* ---- (2) planes: signedDistanceToPlaneEq (Main.hs:1371-1372), first-minimum assignment ------ */
HS_API int32_t hs_plane_assign(hs_ctx* ctx, const hs_cloud* cloud, const float* planes, int32_t K, uint8_t* assign_out,
                               float* resid_out_or_null);
HS_API int32_t hs_plane_assign_dev(hs_ctx* ctx, const hs_cloud* cloud, const float* planes, int32_t K, void* d_assign,
                                   void* d_resid_or_null);
/* cuboid params [x,y,z,a,b,c,q1..q4] (FitCuboidBFGS.hs:99) -> 6 PlaneEq in Float, order +x -x +y -y +z -z
 * (Main.hs:1831-1836 + makePlanesFromCuboid Main.hs:1852-1874).  Host arithmetic, no device needed. */
HS_API int32_t hs_planes_from_cuboid(const double params[10], float planes_out[24]);
/* objective + gradient of the cuboid over a whole cloud: f = sum r_i^2 over nearest-plane residuals,
 * grad = d f / d params, counts = inliers per plane.  Generalises errfun/errfunClosestCenter
 * (FitCuboidBFGS.hs:51-76) from 8 corners to the room cloud. */
HS_API int32_t hs_cuboid_residual_grad(hs_ctx* ctx, const hs_cloud* cloud, const double params[10], double* f,
                                       double grad[10], int64_t counts[6]);
/* raw per-room reductions: rooms are contiguous point ranges [room_offsets[r], room_offsets[r+1]) of `cloud`,
 * each with its own cuboid.  rec_out = nrooms x HS_REC doubles.  Additive across point shards / GPUs. */
HS_API int32_t hs_rooms_cuboid_sums(hs_ctx* ctx, const hs_cloud* cloud, const int64_t* room_offsets, int32_t nrooms,
                                    const double* params, double* rec_out);
/* same, enqueued on the ctx stream with the result left in device memory (nrooms x HS_REC doubles). */
HS_API int32_t hs_rooms_cuboid_sums_async(hs_ctx* ctx, const hs_cloud* cloud, const int64_t* room_offsets,
                                          int32_t nrooms, const double* params, void* d_rec_out);
/* multi-GPU (one process per GPU, one NVSwitch domain, <= 8 ranks): every rank reduces its point range and the records are
 * summed over NVLink peer memory inside the same kernel launch (no NCCL call, no second launch); every rank ends up with
 * the identical total in d_rec_out.  Setup once: each rank creates a mailbox and publishes its 64-byte CUDA IPC handle
 * (any transport: torch.distributed, MPI, a file), then connects with the world's handles in rank order.  All ranks must
 * issue the same sequence of hs_rooms_cuboid_sums_allreduce_async calls. */
HS_API int32_t hs_peer_mailbox_create(hs_ctx* ctx, int32_t rank, int32_t world, uint8_t handle_out[64]);
HS_API int32_t hs_peer_mailbox_connect(hs_ctx* ctx, const uint8_t* handles /* world x 64 bytes */);
HS_API int32_t hs_rooms_cuboid_sums_allreduce_async(hs_ctx* ctx, const hs_cloud* cloud, const int64_t* room_offsets,
                                                    int32_t nrooms, const double* params, void* d_rec_out);
/* Same-process peer group: ctxs[i] (one per device, all in THIS process) becomes rank i of an n-rank group whose mailboxes are
 * addressed directly over NVLink (cudaDeviceEnablePeerAccess).  This is what a single Haskell executable
 * (housescan.cabal:15-43) uses to drive several GPUs; the *_allreduce_async entry points and sessions may then be issued for
 * all ranks from one host thread. */
HS_API int32_t hs_peer_group_create_local(hs_ctx* const* ctxs, int32_t n);

/* ---- evaluation sessions: the optimiser loop's view of the device --------------------------------------------------------
 * FitCuboidBFGS evaluates its objective up to 2000 times per stage (FitCuboidBFGS.hs:184,201,233 `minimize NMSimplex2 1e-8
 * maxIt`); over a room cloud every evaluation is one pass of the reduction kernel, so the fixed cost per launch IS the cost of
 * a fit on several GPUs.  A session launches the kernel once: it stays resident on the SMs, keeps its tile ring primed (the
 * cloud does not change between evaluations), takes parameter sets from a host-mapped command ring and delivers every
 * evaluation's records (summed over the peer group when allreduce != 0) into a host-mapped result ring.  Up to 256
 * evaluations may be posted ahead (line searches, simplex vertices, finite differences); they run back to back.
 *   begin: rooms as in hs_rooms_cuboid_sums (nrooms <= 32); the cloud must stay alive and unchanged until end.
 *   post:  params = count x nrooms x 10 doubles; blocks only while 256 evaluations are in flight.
 *   wait:  blocks until evaluation `seq` (0-based, in posting order) is done; rec_out (may be NULL) = nrooms x HS_REC doubles.
 *          A record stays readable until 256 further evaluations have been posted.
 *   eval:  post one + wait for it.      stop: no more posts; the kernel drains and exits (non-blocking).
 *   end:   stop + wait for the kernel + free.  HS_ENCCL if a peer did not show up within 2 s (its records are NaN).
 * While a session is open every other call on the same ctx fails with HS_EINVAL (the stream is busy with the resident kernel);
 * a session nobody talks to for 20 s shuts itself down.  With a peer group all ranks must post the same evaluations. */
typedef struct hs_eval_session hs_eval_session;
HS_API int32_t hs_eval_session_begin(hs_ctx* ctx, const hs_cloud* cloud, const int64_t* room_offsets, int32_t nrooms,
                                     int32_t allreduce, hs_eval_session** out);
HS_API int32_t hs_eval_session_post(hs_eval_session* s, const double* params, int32_t count);
HS_API int32_t hs_eval_session_wait(hs_eval_session* s, int64_t seq, double* rec_out);
HS_API int32_t hs_eval_session_eval(hs_eval_session* s, const double* params, double* rec_out);
HS_API int64_t hs_eval_session_done(const hs_eval_session* s); /* evaluations finished so far */
/* device clock (%globaltimer, ns) of evaluation `seq`: when the device saw the command, when its records were committed */
HS_API int32_t hs_eval_session_times(const hs_eval_session* s, int64_t seq, uint64_t* seen_ns, uint64_t* done_ns);
HS_API void* hs_eval_session_device_results(const hs_eval_session* s); /* device ring [256][nrooms x HS_REC] doubles */
HS_API int32_t hs_eval_session_stop(hs_eval_session* s);
HS_API int32_t hs_eval_session_end(hs_eval_session* s);
/* The evaluation kernel's static partition of a cloud layout, host only: blocks own contiguous ranges of 4-point groups
 * [first_group[b], first_group[b+1]); a block pays seg_cost groups (0 = default, < 0 = none) for every room boundary inside its
 * range, or ends at the boundary.  block_room_first/last = the rooms with points in the block (last < first: none);
 * room_first_block / room_nblocks = the blocks that deliver a partial record for the room.  Any output but nblocks may be NULL. */
HS_API int32_t hs_eval_plan(int64_t n, const int64_t* room_offsets, int32_t nrooms, int32_t sm_count, int32_t seg_cost,
                            int32_t* nblocks_out, int64_t* block_first_group_out, int32_t* block_room_first_out,
                            int32_t* block_room_last_out, int32_t* room_first_block_out, int32_t* room_nblocks_out);
/* host chain rule: (params, summed record) -> f, grad, counts */
HS_API int32_t hs_cuboid_grad_from_sums(const double params[10], const double rec[HS_REC], double* f, double grad[10],
                                        int64_t counts[6]);
/* generic K planes per room: out[room][k] = count, sum r, sum r^2, sum p (3), sum r p (3), max|r|.
 * Feeds the wall offsets of optimizeRoomPositions (Main.hs:2111-2118, 2188-2190) from points instead of 4+8 corners. */
HS_API int32_t hs_plane_sums(hs_ctx* ctx, const hs_cloud* cloud, const int64_t* room_offsets, int32_t nrooms,
                             const float* planes, int32_t K, double* out);
/* fitPlane's per-point part (Main.hs:1436-1450): mean (Double accumulation, returned also rounded to Float),
 * scatter of Float-centred points in Double: xx,xy,xz,yy,yz,zz. */
HS_API int32_t hs_scatter3x3(hs_ctx* ctx, const hs_cloud* cloud, double mean[3], double scatter[6]);
/* fitPlane complete: smallest eigenvector (host Jacobi) -> PlaneEq; sign as LAPACK leaves it is arbitrary. */
HS_API int32_t hs_fit_plane(hs_ctx* ctx, const hs_cloud* cloud, float plane_out[4]);

/* ---- (3) rigid transforms + export -------------------------------------------------------- */
/* projectRoom's cloud part: translateCloud off . rotateCloudAround zero R (Main.hs:1716,1725-1730).
 * HS_EINVAL if the last column of m is not exactly (0,0,0,1) (the reference pattern-fails).  in == out allowed. */
HS_API int32_t hs_transform(hs_ctx* ctx, const hs_cloud* cloud_in, const float m_rowmajor[16], hs_cloud* cloud_out);
/* rotateCloudAround c R (Main.hs:1657-1659, 1582-1583) and translateCloud (Main.hs:1697-1699) */
HS_API int32_t hs_rotate_around(hs_ctx* ctx, const hs_cloud* cloud_in, const float center[3], const float R_rowmajor[9],
                                hs_cloud* cloud_out);
HS_API int32_t hs_translate(hs_ctx* ctx, const hs_cloud* cloud_in, const float off[3], hs_cloud* cloud_out);
/* pointMean / cloudMean (Main.hs:1596-1605) with Double accumulation + max distance to the Float-rounded mean (Main.hs:1527) */
HS_API int32_t hs_mean_extent(hs_ctx* ctx, const hs_cloud* cloud, double mean[3], float* maxdist);
/* full-resolution export (README.md:16 step 4; replaces the external plyxform / pcl_transform_point_cloud of
 * Main.hs:2311-2313): binary little-endian PLY, float x y z [+ uchar red green blue]. */
HS_API int32_t hs_write_ply(hs_ctx* ctx, const hs_cloud* cloud, const uint8_t* rgb_or_null, const char* path);
/* The same file written by several ranks (per-room transform + export sharded by point range;
 * Main.hs:1716-1730, README.md:16): ONE caller creates the file with its header at the final size (begin), then every rank writes
 * the points [first, first + n) it holds - any order, any process.  Byte-identical to hs_write_ply of the whole cloud. */
HS_API int32_t hs_write_ply_begin(const char* path, int64_t n_total, int32_t has_rgb);
HS_API int32_t hs_write_ply_part_host(const char* path, const float* xyz, const uint8_t* rgb_or_null, int64_t first, int64_t n,
                                      int64_t n_total); /* the same for points that live in host memory */
HS_API int32_t hs_write_ply_part(hs_ctx* ctx, const hs_cloud* cloud, const uint8_t* rgb_or_null, const char* path, int64_t first,
                                 int64_t n_total);
/* ---- room input formats: the step before the hot path (README.md:13-16) ----------------------- */
/* planeEqsFromFile (Main.hs:1379-1389): PCL's planes.txt, one `a b c d` per line meaning ax + by + cz + d = 0; the result is
 * mkPlaneEqABCD a b c (-d).  Parsing stops at the first line that does not match (attoparsec parseOnly); no plane at all is
 * HS_EIO ("Could not load planes").  *n_out receives the number parsed; HS_EINVAL if it exceeds cap. */
HS_API int32_t hs_plane_eqs_from_text(const char* text, int64_t len, float* planes_out, int32_t cap, int32_t* n_out);
HS_API int32_t hs_plane_eqs_from_file(const char* path, float* planes_out, int32_t cap, int32_t* n_out);
/* cloudFromFile (Main.hs:1332-1345) over loadPCDFileXyzFloat / loadPCDFileXyzRgbNormalFloat (:1318-1329): PCD v0.7, DATA ascii,
 * binary or binary_compressed; x y z must be 4-byte floats.  The DATA section goes to the GPU as it lies in the file and is
 * unpacked there.  *colors_out (may be NULL) receives the `ManyColors` cloud r/255 g/255 b/255 when the file has an rgb field,
 * else NULL (`OneColor`).  A file without points is HS_EIO with the reference's message. */
HS_API int32_t hs_cloud_from_pcd(hs_ctx* ctx, const char* path, hs_cloud** cloud_out, hs_cloud** colors_out);
HS_API int32_t hs_pcd_info(const char* path, int64_t* n_points, int32_t* has_rgb, int32_t* data_kind /* 0 ascii 1 binary 2 compressed */);
/* ---- transform export compatibility: the step after the hot path --------------------------------
 * The reference writes .xf files / -matrix strings (Main.hs:2271-2325) for the external plyxform and pcl_transform_point_cloud
 * tools to transform the full-resolution .ply / .pcd of a room.  These entry points close that loop on the GPU: read the file,
 * hs_transform with the parsed matrix, hs_write_ply / hs_write_pcd. */
/* inverse of hs_proj_to_xf / hs_proj_to_string: 16 numbers (blank- or comma-separated) holding the LEFT-multiplicative matrix
 * (Main.hs:2278-2284); the result is roomProj again (row-major, right-multiplied) */
HS_API int32_t hs_transform_from_text(const char* text, int64_t len, float m_rowmajor_out[16]);
/* PLY vertex clouds: format ascii or binary_little_endian, `element vertex` first, float x y z [+ uchar red green blue] */
HS_API int32_t hs_cloud_from_ply(hs_ctx* ctx, const char* path, hs_cloud** cloud_out, hs_cloud** colors_out);
HS_API int32_t hs_ply_info(const char* path, int64_t* n_vertices, int32_t* has_rgb, int32_t* is_ascii); /* header only, no device */
/* binary PCD v0.7, FIELDS x y z [rgb] (what pcl_transform_point_cloud ... -matrix would have written, Main.hs:2305-2313) */
HS_API int32_t hs_write_pcd(hs_ctx* ctx, const hs_cloud* cloud, const uint8_t* rgb_or_null, const char* path);

/* makeInwardFacing (Main.hs:1746-1751): flip (n, d) of every plane unless (roomCenter - planeMean) . n > 0 */
HS_API int32_t hs_make_inward_facing(const float room_center[3], const float* plane_means /* K x 3 */, float* planes_inout /* K x 4 */, int32_t K);
/* loadRoom (Main.hs:1740-1765): dir/cloud_downsampled.pcd, dir/planes.txt, dir/cloud_plane_hull<i>.pcd; planes inward facing */
HS_API int32_t hs_load_room(hs_ctx* ctx, const char* dir, hs_cloud** cloud_out, hs_cloud** colors_out, float* planes_out, int32_t cap, int32_t* K_out);

/* Plane algebra behind the room-editing actions that move a room's planes together with its cloud (host only, Float):
 * rotationBetweenPlaneEqs (Main.hs:1553-1560; row-vector rotation R with n1 .* R parallel to n2, NaN for (anti)parallel normals as
 * in the reference), rotatePlaneEqAround (:1571-1578), translatePlaneEq (:1681-1688).  The cloud side is hs_rotate_around /
 * hs_translate with the same R / offset. */
HS_API int32_t hs_rotation_between_plane_eqs(const float plane1[4], const float plane2[4], float R_rowmajor_out[9]);
HS_API int32_t hs_rotate_plane_eq_around(const float center[3], const float R_rowmajor[9], const float plane_in[4], float plane_out[4]);
HS_API int32_t hs_translate_plane_eq(const float offset[3], const float plane_in[4], float plane_out[4]);
/* roomProj bookkeeping (row-major 4x4, right-multiplied, Float): `a .*. b` (projectRoom, Main.hs:1720), `translate4 off proj`
 * (translateRoom, :1708), `translate4 c . (.*. linear R) . translate4 (neg c) $ proj` (rotateRoomAround, :1674) */
HS_API int32_t hs_proj_compose(const float a[16], const float b[16], float out[16]);
HS_API int32_t hs_proj_translate(const float proj[16], const float offset[3], float out[16]);
HS_API int32_t hs_proj_rotate_around(const float proj[16], const float center[3], const float R_rowmajor[9], float out[16]);
/* planeCorner (Main.hs:1413-1430): the corner where three planes meet (3x3 solve in Double, result in Float);
 * HS_ESINGULAR = Nothing when the system has an exactly zero pivot (parallel planes) */
HS_API int32_t hs_plane_corner(const float plane1[4], const float plane2[4], const float plane3[4], float corner_out[3]);
/* roomProjectionToString / roomProjectionToXfFormat (Main.hs:2271-2302); buf receives a NUL-terminated string */
HS_API int32_t hs_proj_to_string(const float m_rowmajor[16], char* buf, int32_t buflen);
HS_API int32_t hs_proj_to_xf(const float m_rowmajor[16], char* buf, int32_t buflen);

/* ---- (4) connected components: GroupConnectedComponents.groupCCContiguous on dense ids
 *      (GroupConnectedComponents.hs:39-54); label = minimum vertex id of the component ------ */
HS_API int32_t hs_cc_label(hs_ctx* ctx, const uint32_t* src, const uint32_t* dst, int64_t E, uint32_t N,
                           uint32_t* label_out);
HS_API int32_t hs_cc_label_dev(hs_ctx* ctx, const void* d_src, const void* d_dst, int64_t E, uint32_t N, void* d_label);

/* ---- VectorUtil.kthLargestBy / kthSmallestBy on Float keys (VectorUtil.hs:11-19), removeCeiling (Main.hs:2643-2664) */
/* keys read from the cloud at component `axis` (0,1,2).  k is 1-based; k<1 or k>n -> HS_EINVAL with the reference's message. */
HS_API int32_t hs_kth_largest(hs_ctx* ctx, const hs_cloud* cloud, int32_t axis, int64_t k, float* out);
HS_API int32_t hs_kth_smallest(hs_ctx* ctx, const hs_cloud* cloud, int32_t axis, int64_t k, float* out);
/* order-preserving V.filter ((<= limit) . component axis).  `colors` (the ManyColors Vector Vec3 of Main.hs:112-115, same
 * length as the cloud) is filtered by the same predicate on the POINTS (V.ifilter, Main.hs:2664) when given. */
/* Sharded k-th (point ranges on several GPUs).  One call = one MSB radix pass over THIS rank's points:
 * hist_out[2048] is the local histogram of the pass's digit (pass 0: bits 31..21, 1: bits 20..10, 2: bits 9..0 of the order-
 * preserving key image) over the keys with (key & mask) == prefix.  The caller sums the histograms of all ranks (one all-reduce
 * of 2048 counters per pass), picks the digit that holds the k-th key, extends (prefix, mask) and calls the next pass; after
 * pass 2 the prefix is the key of the k-th value.  Same semantics as kthLargestBy / kthSmallestBy (VectorUtil.hs:11-19). */
HS_API int32_t hs_kth_shard_pass(hs_ctx* ctx, const hs_cloud* cloud, int32_t axis, int32_t pass, uint32_t prefix, uint32_t mask, uint32_t* hist_out);
HS_API uint32_t hs_kth_key_of_float(float v);   /* order-preserving uint image of a Float key */
HS_API float hs_kth_float_of_key(uint32_t key);
HS_API int32_t hs_filter_le(hs_ctx* ctx, const hs_cloud* cloud, int32_t axis, float limit, const hs_cloud* colors_or_null,
                            hs_cloud* cloud_out, hs_cloud* colors_out_or_null, int64_t* n_out);
HS_API int32_t hs_remove_ceiling(hs_ctx* ctx, const hs_cloud* cloud, const hs_cloud* colors_or_null, hs_cloud* cloud_out,
                                 hs_cloud* colors_out_or_null, int64_t* n_out, float* y_limit);

/* ---- host-side mirrors of the Haskell modules (C++ in housescan_b200/host, exported flat for FFI/tests) ---- */
/* FitCuboidBFGS.cuboidFromParams / errfun (FitCuboidBFGS.hs:51-112), 8 corners x 3 doubles */
HS_API int32_t hs_cuboid_from_params(const double params[10], double corners_out[24]);
HS_API double hs_errfun(const double corners[24], const double params[10]);
HS_API double hs_errfun_closest(const double* pts, int32_t npts, const double params[10]);
HS_API int32_t hs_guess_dims(const double corners[24], double out[3]);
/* fitCuboid / fitCuboidFromCenter / fitCuboidFromCenterFirst (FitCuboidBFGS.hs:172-233): Nelder-Mead simplex
 * (GSL nmsimplex2 semantics), tol 1e-8, maxit 2000.  path_out may be NULL; else rows x (3 + nparams) doubles, up to path_cap rows. */
HS_API int32_t hs_fit_cuboid(const double corners[24], int32_t variant /*0 fitCuboid,1 FromCenter,2 FromCenterFirst*/,
                             double params_out[10], int32_t* steps, double* err, double* path_out, int32_t path_cap);
/* north-star addition: BFGS over the whole cloud using hs_cuboid_residual_grad on the GPU */
HS_API int32_t hs_fit_cuboid_cloud_bfgs(hs_ctx* ctx, const hs_cloud* cloud, const double init[10], int32_t max_iter,
                                        double gtol, double params_out[10], double* f_out, int32_t* iters,
                                        int32_t* evals);
/* The same BFGS (host/hs_host.cpp) over a caller-supplied objective: eval(user, x, &f, g) returns 0 on success.  What a host
 * program uses to minimise any of the records' objectives with its own chain rule (FitCuboidBFGS.hs:172-252 drives GSL the same
 * way through hmatrix-gsl's callbacks), and what the parity tests use to run the identical optimiser over the oracle's objective. */
typedef int32_t (*hs_objective_fn)(void* user, const double* x, double* f, double* grad);
HS_API int32_t hs_bfgs_minimize(hs_objective_fn eval, void* user, const double* x0, int32_t n, int32_t max_iter, double gtol,
                                double* x_out, double* f_out, int32_t* iters, int32_t* evals);
/* The reference's own optimiser over a caller-supplied objective: Nelder-Mead with GSL nmsimplex2's update rules as hmatrix's
 * `minimize NMSimplex2 eps maxit x0 f step` drives it (FitCuboidBFGS.hs:184,201,233).  value(user, x) returns the objective. */
typedef double (*hs_value_fn)(void* user, const double* x);
HS_API int32_t hs_nm_minimize(hs_value_fn value, void* user, const double* x0, const double* step, int32_t n, double eps,
                              int32_t maxit, double* x_out, double* f_out, int32_t* iters, int32_t* evals);
/* ... and over a device-resident room cloud: f(params) = sum of squared distances to the nearest wall (the objective of
 * hs_cuboid_residual_grad), evaluated by ONE resident kernel (an evaluation session); the independent evaluations of the initial
 * simplex and of every shrink step are posted together and run back to back.  step[10] = initial simplex sizes. */
HS_API int32_t hs_fit_cuboid_cloud_nm(hs_ctx* ctx, const hs_cloud* cloud, const double init[10], const double step[10], double eps,
                                      int32_t maxit, double params_out[10], double* f_out, int32_t* iters, int32_t* evals);
/* TranslationOptimizer.lstSqDistancesI (TranslationOptimizer.hs:48-72) on bijected indices; HS_ESINGULAR => Nothing */
HS_API int32_t hs_lstsq_distances(const int32_t* i_idx, const int32_t* j_idx, const double* d, int32_t m, int32_t n_nodes,
                                  double* pos_out, double* rmse_out);
/* GroupConnectedComponents.groupConnectedComponents (GroupConnectedComponents.hs:16-32) on already-bijected edges:
 * comp_out[e] = component index (ascending min vertex), order_out = edge indices grouped by component in the
 * reference's output order (reverse input order inside a component).  Labels computed on the GPU. */
HS_API int32_t hs_group_cc(hs_ctx* ctx, const uint32_t* src, const uint32_t* dst, int64_t E, uint32_t N,
                           int32_t* comp_out, int64_t* order_out, int32_t* ncomp_out);

HS_API const char* hs_version(void);

#ifdef __cplusplus
}
#endif
#endif
