-- Binding of include/housescan_b200.h for the HouseScan executable (see INTEGRATION.md).  Not compiled in this repository's
-- image (no GHC); the identical symbols are exercised through ctypes by the test-suite.
{-# LANGUAGE ForeignFunctionInterface, EmptyDataDecls #-}
module HouseScanB200.FFI where

import Data.Int (Int32, Int64)
import Data.Word (Word8, Word16, Word32, Word64)
import Foreign.C.String (CString)
import Foreign.C.Types (CDouble(..), CFloat(..))
import Foreign.Ptr (Ptr)

data HsCtx      -- hs_ctx:   one CUDA device + stream; one in-flight call per ctx
data HsCloud    -- hs_cloud: device-resident `Vector Vec3` (12 B/point AoS, Main.hs:39-42,120)

-- context -----------------------------------------------------------------------------------------------------
foreign import ccall safe "hs_ctx_create"   c_ctx_create   :: Int32 -> Ptr (Ptr HsCtx) -> IO Int32
foreign import ccall safe "hs_ctx_destroy"  c_ctx_destroy  :: Ptr HsCtx -> IO Int32
foreign import ccall unsafe "hs_last_error" c_last_error   :: Ptr HsCtx -> IO CString
-- clouds (Cloud.cloudPoints, Main.hs:117-121) ---------------------------------------------------------------------
foreign import ccall safe "hs_cloud_upload"   c_cloud_upload   :: Ptr HsCtx -> Ptr CFloat -> Int64 -> Ptr (Ptr HsCloud) -> IO Int32
foreign import ccall safe "hs_cloud_alloc"    c_cloud_alloc    :: Ptr HsCtx -> Int64 -> Ptr (Ptr HsCloud) -> IO Int32
foreign import ccall safe "hs_cloud_download" c_cloud_download :: Ptr HsCtx -> Ptr HsCloud -> Ptr CFloat -> IO Int32
foreign import ccall unsafe "hs_cloud_size"   c_cloud_size     :: Ptr HsCloud -> IO Int64
foreign import ccall safe "hs_cloud_free"     c_cloud_free     :: Ptr HsCtx -> Ptr HsCloud -> IO Int32
-- (1) depth frames: HoniHelper.takeDepthSnapshot's (Vector Word16,(w,h)) -> addDevicePointCloud (Main.hs:1296-1313)
foreign import ccall safe "hs_backproject_ref"
  c_backproject_ref :: Ptr HsCtx -> Ptr Word16 -> Int32 -> Int32 -> Ptr CFloat -> Ptr Word8 -> Ptr Int64 -> IO Int32
foreign import ccall safe "hs_backproject_reduce6x6"
  c_backproject_reduce6x6 :: Ptr HsCtx -> Ptr Word16 -> Int64 -> Int32 -> Int32 -> Ptr CFloat -> Ptr CFloat
                          -> Ptr CFloat -> Int32 -> Ptr CDouble -> IO Int32
-- (2) planes: signedDistanceToPlaneEq (Main.hs:1371-1372), cuboid objective (FitCuboidBFGS.hs:51-76 generalised)
foreign import ccall safe "hs_plane_assign"
  c_plane_assign :: Ptr HsCtx -> Ptr HsCloud -> Ptr CFloat -> Int32 -> Ptr Word8 -> Ptr CFloat -> IO Int32
foreign import ccall safe "hs_cuboid_residual_grad"
  c_cuboid_residual_grad :: Ptr HsCtx -> Ptr HsCloud -> Ptr CDouble -> Ptr CDouble -> Ptr CDouble -> Ptr Int64 -> IO Int32
foreign import ccall safe "hs_rooms_cuboid_sums"
  c_rooms_cuboid_sums :: Ptr HsCtx -> Ptr HsCloud -> Ptr Int64 -> Int32 -> Ptr CDouble -> Ptr CDouble -> IO Int32
foreign import ccall safe "hs_plane_sums"
  c_plane_sums :: Ptr HsCtx -> Ptr HsCloud -> Ptr Int64 -> Int32 -> Ptr CFloat -> Int32 -> Ptr CDouble -> IO Int32
foreign import ccall safe "hs_fit_plane"      c_fit_plane   :: Ptr HsCtx -> Ptr HsCloud -> Ptr CFloat -> IO Int32
foreign import ccall safe "hs_fit_cuboid_cloud_bfgs"
  c_fit_cuboid_cloud_bfgs :: Ptr HsCtx -> Ptr HsCloud -> Ptr CDouble -> Int32 -> CDouble -> Ptr CDouble -> Ptr CDouble
                          -> Ptr Int32 -> Ptr Int32 -> IO Int32
-- (3) transforms + export: projectRoom (Main.hs:1716-1730), rotateCloudAround (:1657-1659), translateCloud (:1697-1699)
foreign import ccall safe "hs_transform"     c_transform     :: Ptr HsCtx -> Ptr HsCloud -> Ptr CFloat -> Ptr HsCloud -> IO Int32
foreign import ccall safe "hs_rotate_around" c_rotate_around :: Ptr HsCtx -> Ptr HsCloud -> Ptr CFloat -> Ptr CFloat -> Ptr HsCloud -> IO Int32
foreign import ccall safe "hs_translate"     c_translate     :: Ptr HsCtx -> Ptr HsCloud -> Ptr CFloat -> Ptr HsCloud -> IO Int32
foreign import ccall safe "hs_mean_extent"   c_mean_extent   :: Ptr HsCtx -> Ptr HsCloud -> Ptr CDouble -> Ptr CFloat -> IO Int32
foreign import ccall safe "hs_write_ply"     c_write_ply     :: Ptr HsCtx -> Ptr HsCloud -> Ptr Word8 -> CString -> IO Int32
-- (4) connected components on bijected ids (GroupConnectedComponents.hs:39-54)
foreign import ccall safe "hs_cc_label"
  c_cc_label :: Ptr HsCtx -> Ptr Word32 -> Ptr Word32 -> Int64 -> Word32 -> Ptr Word32 -> IO Int32
foreign import ccall safe "hs_group_cc"
  c_group_cc :: Ptr HsCtx -> Ptr Word32 -> Ptr Word32 -> Int64 -> Word32 -> Ptr Int32 -> Ptr Int64 -> Ptr Int32 -> IO Int32
-- VectorUtil.kthLargestBy / removeCeiling (VectorUtil.hs:11-19, Main.hs:2643-2664)
foreign import ccall safe "hs_kth_largest"    c_kth_largest    :: Ptr HsCtx -> Ptr HsCloud -> Int32 -> Int64 -> Ptr CFloat -> IO Int32
foreign import ccall safe "hs_remove_ceiling"
  c_remove_ceiling :: Ptr HsCtx -> Ptr HsCloud -> Ptr HsCloud -> Ptr HsCloud -> Ptr HsCloud -> Ptr Int64 -> Ptr CFloat -> IO Int32
-- host-only mirrors (no device): TranslationOptimizer.lstSqDistancesI, FitCuboidBFGS.fitCuboid*
foreign import ccall unsafe "hs_lstsq_distances"
  c_lstsq_distances :: Ptr Int32 -> Ptr Int32 -> Ptr CDouble -> Int32 -> Int32 -> Ptr CDouble -> Ptr CDouble -> IO Int32
foreign import ccall safe "hs_fit_cuboid"
  c_fit_cuboid :: Ptr CDouble -> Int32 -> Ptr CDouble -> Ptr Int32 -> Ptr CDouble -> Ptr CDouble -> Int32 -> IO Int32

-- multi-GPU peer group (INTEGRATION.md section 5)
foreign import ccall safe "hs_peer_mailbox_create"  c_peer_mailbox_create  :: Ptr HsCtx -> Int32 -> Int32 -> Ptr Word8 -> IO Int32
foreign import ccall safe "hs_peer_mailbox_connect" c_peer_mailbox_connect :: Ptr HsCtx -> Ptr Word8 -> IO Int32
foreign import ccall safe "hs_rooms_cuboid_sums_allreduce_async"
  c_rooms_cuboid_sums_allreduce :: Ptr HsCtx -> Ptr HsCloud -> Ptr Int64 -> Int32 -> Ptr CDouble -> Ptr () -> IO Int32

-- one executable, several GPUs: ctxs of this process become ranks 0..n-1 (housescan.cabal:15-43 builds ONE executable)
foreign import ccall safe "hs_peer_group_create_local" c_peer_group_create_local :: Ptr (Ptr HsCtx) -> Int32 -> IO Int32

-- evaluation sessions: the optimiser loop of FitCuboidBFGS.hs:184,201,233 against ONE resident kernel (no launch per evaluation)
data HsEvalSession
foreign import ccall safe "hs_eval_session_begin"
  c_eval_session_begin :: Ptr HsCtx -> Ptr HsCloud -> Ptr Int64 -> Int32 -> Int32 -> Ptr (Ptr HsEvalSession) -> IO Int32
foreign import ccall safe "hs_eval_session_post"  c_eval_session_post  :: Ptr HsEvalSession -> Ptr CDouble -> Int32 -> IO Int32
foreign import ccall safe "hs_eval_session_wait"  c_eval_session_wait  :: Ptr HsEvalSession -> Int64 -> Ptr CDouble -> IO Int32
foreign import ccall safe "hs_eval_session_eval"  c_eval_session_eval  :: Ptr HsEvalSession -> Ptr CDouble -> Ptr CDouble -> IO Int32
foreign import ccall unsafe "hs_eval_session_done" c_eval_session_done :: Ptr HsEvalSession -> IO Int64
foreign import ccall unsafe "hs_eval_session_times" c_eval_session_times :: Ptr HsEvalSession -> Int64 -> Ptr Word64 -> Ptr Word64 -> IO Int32
foreign import ccall unsafe "hs_eval_session_device_results" c_eval_session_device_results :: Ptr HsEvalSession -> IO (Ptr ())
foreign import ccall safe "hs_eval_session_stop"  c_eval_session_stop  :: Ptr HsEvalSession -> IO Int32
foreign import ccall safe "hs_eval_session_end"   c_eval_session_end   :: Ptr HsEvalSession -> IO Int32

-- room input formats (SURVEY.md 8f rank 1): planeEqsFromFile Main.hs:1379-1389, cloudFromFile :1332-1345, loadRoom :1740-1765
foreign import ccall unsafe "hs_plane_eqs_from_text"
  c_plane_eqs_from_text :: CString -> Int64 -> Ptr CFloat -> Int32 -> Ptr Int32 -> IO Int32
foreign import ccall safe "hs_plane_eqs_from_file"
  c_plane_eqs_from_file :: CString -> Ptr CFloat -> Int32 -> Ptr Int32 -> IO Int32
foreign import ccall safe "hs_cloud_from_pcd"
  c_cloud_from_pcd :: Ptr HsCtx -> CString -> Ptr (Ptr HsCloud) -> Ptr (Ptr HsCloud) -> IO Int32
foreign import ccall safe "hs_pcd_info"
  c_pcd_info :: CString -> Ptr Int64 -> Ptr Int32 -> Ptr Int32 -> IO Int32
foreign import ccall unsafe "hs_make_inward_facing"
  c_make_inward_facing :: Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> Int32 -> IO Int32
foreign import ccall safe "hs_load_room"
  c_load_room :: Ptr HsCtx -> CString -> Ptr (Ptr HsCloud) -> Ptr (Ptr HsCloud) -> Ptr CFloat -> Int32 -> Ptr Int32 -> IO Int32

-- transform export compatibility (SURVEY.md 8f rank 2): replaces the external plyxform / pcl_transform_point_cloud step (Main.hs:2305-2325)
foreign import ccall unsafe "hs_transform_from_text" c_transform_from_text :: CString -> Int64 -> Ptr CFloat -> IO Int32
foreign import ccall safe "hs_cloud_from_ply"
  c_cloud_from_ply :: Ptr HsCtx -> CString -> Ptr (Ptr HsCloud) -> Ptr (Ptr HsCloud) -> IO Int32
foreign import ccall safe "hs_write_pcd" c_write_pcd :: Ptr HsCtx -> Ptr HsCloud -> Ptr Word8 -> CString -> IO Int32
-- sharded k-th (point ranges on several GPUs, SURVEY.md 8e): one radix pass per call, histograms summed by the caller
foreign import ccall safe "hs_kth_shard_pass"
  c_kth_shard_pass :: Ptr HsCtx -> Ptr HsCloud -> Int32 -> Int32 -> Word32 -> Word32 -> Ptr Word32 -> IO Int32
foreign import ccall unsafe "hs_kth_float_of_key" c_kth_float_of_key :: Word32 -> CFloat
-- plane algebra of the room-editing actions (Main.hs:1553-1578, :1681-1688), host only
foreign import ccall unsafe "hs_rotation_between_plane_eqs" c_rotation_between_plane_eqs :: Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> IO Int32
foreign import ccall unsafe "hs_rotate_plane_eq_around" c_rotate_plane_eq_around :: Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> IO Int32
foreign import ccall unsafe "hs_translate_plane_eq" c_translate_plane_eq :: Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> IO Int32
foreign import ccall unsafe "hs_plane_corner" c_plane_corner :: Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> IO Int32
-- roomProj bookkeeping (Main.hs:1674, :1708, :1720)
foreign import ccall unsafe "hs_proj_compose" c_proj_compose :: Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> IO Int32
foreign import ccall unsafe "hs_proj_translate" c_proj_translate :: Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> IO Int32
foreign import ccall unsafe "hs_proj_rotate_around" c_proj_rotate_around :: Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> IO Int32
foreign import ccall safe "hs_ply_info" c_ply_info :: CString -> Ptr Int64 -> Ptr Int32 -> Ptr Int32 -> IO Int32

-- host-only mirrors used by the FitCuboidBFGS shim (FitCuboidBFGS.hs:51-131, :247-252)
foreign import ccall unsafe "hs_cuboid_from_params" c_cuboid_from_params :: Ptr CDouble -> Ptr CDouble -> IO Int32
foreign import ccall unsafe "hs_errfun"             c_errfun             :: Ptr CDouble -> Ptr CDouble -> IO CDouble
foreign import ccall unsafe "hs_guess_dims"         c_guess_dims         :: Ptr CDouble -> Ptr CDouble -> IO Int32
foreign import ccall unsafe "hs_cuboid_grad_from_sums"
  c_cuboid_grad_from_sums :: Ptr CDouble -> Ptr CDouble -> Ptr CDouble -> Ptr CDouble -> Ptr Int64 -> IO Int32
foreign import ccall unsafe "hs_planes_from_cuboid" c_planes_from_cuboid :: Ptr CDouble -> Ptr CFloat -> IO Int32
foreign import ccall safe "hs_kth_smallest"         c_kth_smallest       :: Ptr HsCtx -> Ptr HsCloud -> Int32 -> Int64 -> Ptr CFloat -> IO Int32
-- sharded full-resolution export (SURVEY.md 8e row 3)
foreign import ccall safe "hs_write_ply_begin"     c_write_ply_begin     :: CString -> Int64 -> Int32 -> IO Int32
foreign import ccall safe "hs_write_ply_part"      c_write_ply_part      :: Ptr HsCtx -> Ptr HsCloud -> Ptr Word8 -> CString -> Int64 -> Int64 -> IO Int32
foreign import ccall safe "hs_write_ply_part_host" c_write_ply_part_host :: CString -> Ptr CFloat -> Ptr Word8 -> Int64 -> Int64 -> Int64 -> IO Int32
-- the evaluation kernel's static partition of a cloud layout (host only; diagnostics)
foreign import ccall unsafe "hs_eval_plan"
  c_eval_plan :: Int64 -> Ptr Int64 -> Int32 -> Int32 -> Int32 -> Ptr Int32 -> Ptr Int64 -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> IO Int32
-- the reference's optimiser (NMSimplex2) over a device-resident cloud, driven through an evaluation session
foreign import ccall safe "hs_fit_cuboid_cloud_nm"
  c_fit_cuboid_cloud_nm :: Ptr HsCtx -> Ptr HsCloud -> Ptr CDouble -> Ptr CDouble -> CDouble -> Int32 -> Ptr CDouble -> Ptr CDouble -> Ptr Int32 -> Ptr Int32 -> IO Int32
