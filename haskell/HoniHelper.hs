-- Drop-in for housescan/HoniHelper.hs: export list verbatim (HoniHelper.hs:3-6) + back-projection of the snapshot on the GPU.
-- takeDepthSnapshot / withHoni stay what they are in the reference (they talk to OpenNI through the `honi` package, which is not
-- part of this path); the frame they deliver - (Vector Word16, (w, h)) - goes to the device as it comes.
-- Not compiled in this repository's image (no GHC).
{-# LANGUAGE NamedFieldPuns, LambdaCase #-}

module HoniHelper
  ( takeDepthSnapshot
  , withHoni
  -- additive: Main.addDevicePointCloud's per-pixel work (Main.hs:1296-1313) on the GPU
  , backProject
  , backProjectReduce6x6
  ) where

import qualified Data.Vector.Storable as V
import qualified Data.Vector.Storable.Mutable as VM
import           Data.Vect.Float (Vec3)
import           Data.Word (Word8, Word16)
import           Foreign.Marshal.Alloc (alloca)
import           Foreign.Marshal.Array (allocaArray)
import           Foreign.Ptr (castPtr, nullPtr)
import           Foreign.Storable (peek)

import           HouseScanB200.Device
import           HouseScanB200.FFI
import qualified HoniHelper.OpenNI as OpenNI   -- the reference's own HoniHelper.hs:20-56, moved aside unchanged by the maintainer

takeDepthSnapshot :: IO (Either String (V.Vector Word16, (Int, Int)))
takeDepthSnapshot = OpenNI.takeDepthSnapshot

withHoni :: IO a -> IO a
withHoni = OpenNI.withHoni

-- | depth frame -> (points of the pixels with d /= 0 in raster order, the d /= 0 mask): x = column, y = row, z = depth, scaled by
-- the reference's true divisions (Main.hs:1296-1302, scalePoints); bit-exact with the Haskell list code.
backProject :: Ctx -> (V.Vector Word16, (Int, Int)) -> IO (V.Vector Vec3, V.Vector Word8)
backProject ctx (depth, (w, h)) = V.unsafeWith depth $ \pd -> do
  out <- VM.new (w * h); mask <- VM.new (w * h)
  n <- alloca $ \pn -> do
    check ctx =<< withCtxPtr ctx (\c -> VM.unsafeWith out $ \po -> VM.unsafeWith mask $ \pm ->
      c_backproject_ref c pd (fromIntegral w) (fromIntegral h) (castPtr po) pm pn)
    peek pn
  (,) <$> (V.take (fromIntegral n) <$> V.unsafeFreeze out) <*> V.unsafeFreeze mask

-- | a replayed depth stream: per frame the point-to-plane 6x6 normal equations (29 doubles: 21 + 6 + residual + count) against the
-- given planes, fused with the back-projection - the frames never exist as point clouds.  Intrinsics / poses Nothing = the
-- reference's pixel-coordinate points and the identity pose.
backProjectReduce6x6 :: Ctx -> V.Vector Word16 -> Int -> (Int, Int) -> V.Vector Float -> Int -> IO [[Double]]
backProjectReduce6x6 ctx frames nframes (w, h) planes k = V.unsafeWith frames $ \pf -> V.unsafeWith planes $ \pp ->
  allocaArray (nframes * 29) $ \pout -> do
    check ctx =<< withCtxPtr ctx (\c ->
      c_backproject_reduce6x6 c pf (fromIntegral nframes) (fromIntegral w) (fromIntegral h) nullPtr nullPtr (castPtr pp) (fromIntegral k) pout)
    chunk29 <$> peekDoubles (nframes * 29) pout
  where chunk29 [] = []
        chunk29 xs = let (a, b) = splitAt 29 xs in a : chunk29 b
