-- Drop-in for housescan/TranslationOptimizer.hs: export list verbatim (TranslationOptimizer.hs:3-5) + the device-side per-room sums.
-- lstSqDistances keeps its type; lstSqDistancesI's QR solve (TranslationOptimizer.hs:48-72) is host code of the library with the
-- reference's quirks (first node pinned to 0, RMSE over the equations); the per-room sums that feed optimizeRoomPositions
-- (Main.hs:2039-2168) come from ONE launch over the apartment cloud.  Not compiled in this repository's image (no GHC).
{-# LANGUAGE MultiWayIf #-}

module TranslationOptimizer
  ( lstSqDistances
  -- additive
  , RMSE
  , roomsCuboidSums
  ) where

import Data.Int (Int32, Int64)
import Data.Map (Map)
import qualified Data.Map as Map
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (allocaArray, withArray, withArrayLen)
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import Bijection (biject)
import HouseScanB200.Device
import HouseScanB200.FFI

type RMSE = Double

-- | distances between pairs of nodes -> 1-D positions (first node at 0) and the RMSE of the fit; Nothing if the system is singular
lstSqDistances :: (Ord a) => Map (a, a) Double -> Maybe (Map a Double, RMSE)
lstSqDistances m
  | Map.null m = Nothing
  | otherwise  = unsafePerformIO $
      withArray (map (fromIntegral . to . fst) ks :: [Int32]) $ \pi ->
      withArray (map (fromIntegral . to . snd) ks :: [Int32]) $ \pj ->
      withDoubles (Map.elems m) $ \pd -> allocaArray n $ \ppos -> alloca $ \prmse -> do
        rc <- c_lstsq_distances pi pj pd (fromIntegral (length ks)) (fromIntegral n) ppos prmse
        if | rc == 5   -> return Nothing                                   -- HS_ESINGULAR: linearSolveLS had no unique answer
           | rc /= 0   -> error "TranslationOptimizer: hs_lstsq_distances failed"
           | otherwise -> do pos <- peekDoubles n ppos
                             rmse <- realToFrac <$> peek prmse
                             return (Just (Map.fromList [ (from i, p) | (i, p) <- zip [0 ..] pos ], rmse))
  where ks         = Map.keys m
        nodes      = concat [ [a, b] | (a, b) <- ks ]
        (to, from) = biject nodes                                           -- Bijection.hs:16-25, unchanged
        n          = length (Map.toList (Map.fromList [ (x, ()) | x <- nodes ]))

-- | per-room cuboid objective / gradient records (24 doubles each) of an apartment cloud whose rooms are the point ranges
-- [offsets !! r, offsets !! (r+1)): the sums TranslationOptimizer's callers need for every room, from ONE pass over the cloud.
roomsCuboidSums :: Ctx -> DeviceCloud -> [Int] -> [[Double]] -> IO [[Double]]
roomsCuboidSums ctx dc offsets params =
  withArrayLen (map fromIntegral offsets :: [Int64]) $ \len po -> withDoubles (concat params) $ \pp ->
    allocaArray ((len - 1) * 24) $ \pr -> do
      check ctx =<< withCtxPtr ctx (\c -> withCloudPtr dc $ \pc -> c_rooms_cuboid_sums c pc po (fromIntegral (len - 1)) pp pr)
      chunk24 <$> peekDoubles ((len - 1) * 24) pr
  where chunk24 [] = []
        chunk24 xs = let (a, b) = splitAt 24 xs in a : chunk24 b
