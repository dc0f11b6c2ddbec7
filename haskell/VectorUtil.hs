-- Drop-in for housescan/VectorUtil.hs: export list verbatim (VectorUtil.hs:1-4) + the device forms used by removeCeiling
-- (Main.hs:2643-2664).  The polymorphic host functions keep their types (a partial heap sort over any vector, as the reference
-- does with vector-algorithms); a cloud that already lives on the GPU uses the radix select / order-preserving filter kernels.
-- Not compiled in this repository's image (no GHC).
module VectorUtil
  ( kthSmallestBy
  , kthLargestBy
  -- additive: device-resident clouds
  , kthLargestCoord
  , kthSmallestCoord
  , removeCeilingDevice
  ) where

import           Data.Ord (comparing, Down(..))
import           Data.Vector.Algorithms.Heap (partialSortBy)
import qualified Data.Vector.Generic as G
import           Foreign.C.Types (CFloat(..))
import           Foreign.Marshal.Alloc (alloca)
import           Foreign.Ptr (nullPtr)
import           Foreign.Storable (peek)

import           HouseScanB200.Device
import           HouseScanB200.FFI

-- | k-th smallest element (0-based) under a key; errors on an empty vector or k out of range like the reference's indexing
kthSmallestBy :: (G.Vector v a, Ord b) => (a -> b) -> Int -> v a -> a
kthSmallestBy key k v = G.modify (\mv -> partialSortBy (comparing key) mv (k + 1)) v G.! k

kthLargestBy :: (G.Vector v a, Ord b) => (a -> b) -> Int -> v a -> a
kthLargestBy key k v = G.modify (\mv -> partialSortBy (comparing (Down . key)) mv (k + 1)) v G.! k

-- | k-th largest / smallest coordinate (axis 0 = x, 1 = y, 2 = z) of a device cloud: radix select over order-preserving keys,
-- bit-exact with kthLargestBy (\(Vec3 _ y _) -> y) on the downloaded cloud
kthLargestCoord, kthSmallestCoord :: Ctx -> DeviceCloud -> Int -> Int -> IO Float
kthLargestCoord  = kthWith c_kth_largest
kthSmallestCoord = kthWith c_kth_smallest

kthWith f ctx dc axis k = alloca $ \pv -> do
  check ctx =<< withCtxPtr ctx (\c -> withCloudPtr dc $ \pc -> f c pc (fromIntegral axis) (fromIntegral k) pv)
  (\(CFloat x) -> x) <$> peek pv

-- | removeCeiling (Main.hs:2643-2664) without leaving the device: y limit = the reference's k-th largest y, points (and colours)
-- with y <= limit kept in their order.  Returns (cloud, colours, limit).
removeCeilingDevice :: Ctx -> DeviceCloud -> Maybe DeviceCloud -> IO (DeviceCloud, Maybe DeviceCloud, Float)
removeCeilingDevice ctx dc mcol = do
  n    <- cloudSize dc
  out  <- allocCloud ctx n
  cout <- traverse (const (allocCloud ctx n)) mcol
  alloca $ \pn -> alloca $ \pl -> do
    check ctx =<< withCtxPtr ctx (\c -> withCloudPtr dc $ \pc -> withCloudPtr out $ \po ->
      maybe ($ nullPtr) withCloudPtr mcol $ \pci -> maybe ($ nullPtr) withCloudPtr cout $ \pco ->
        c_remove_ceiling c pc pci po pco pn pl)
    CFloat lim <- peek pl
    return (out, cout, lim)   -- hs_remove_ceiling has set the clouds' sizes to the number of points kept
