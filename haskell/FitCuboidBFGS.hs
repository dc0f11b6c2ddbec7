-- Drop-in for housescan/FitCuboidBFGS.hs: the export list of the reference (FitCuboidBFGS.hs:3-12) is kept verbatim, the additive
-- exports below it are the device path.  Bodies call libhousescan_b200.so (HouseScanB200.FFI); the arithmetic the reference does
-- in Haskell over 8 corners is host code of the library with the same results (tests/test_host_logic.py pins it to the
-- reference's own QuickCheck property, example box and `show` literals); the objective over a room cloud runs on the GPU.
-- Not compiled in this repository's image (no GHC).
{-# LANGUAGE ScopedTypeVariables #-}

module FitCuboidBFGS
  ( Cuboid
  , errfun
  , cuboidFromParams
  , fitCuboid
  , fitCuboidFromCenter
  , fitCuboidFromCenterFirst
  , fitCuboidFromCenterFirstError
  , main
  -- additive: the objective and the fit over a device-resident room cloud
  , cuboidResidualGrad
  , fitCuboidToCloudBFGS
  , fitCuboidToCloudNM
  , withCloudObjective
  ) where

import Control.Exception (bracket)
import Data.Vect.Double (Vec3(..))
import Foreign.C.Types (CDouble(..))
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (allocaArray, peekArray, withArray)
import Foreign.Ptr (Ptr, nullPtr)
import Foreign.Storable (peek)
import Numeric.LinearAlgebra (Matrix, fromLists)
import System.IO.Unsafe (unsafePerformIO)

import HouseScanB200.Device
import HouseScanB200.FFI

type Cuboid = [Vec3]   -- 8 corners (FitCuboidBFGS.hs:27)

withCorners :: Cuboid -> (Ptr CDouble -> IO a) -> IO a
withCorners ps k
  | length ps /= 8 = error "FitCuboidBFGS: a Cuboid has 8 corners"
  | otherwise      = withArray (concat [ map realToFrac [x, y, z] | Vec3 x y z <- ps ]) k

-- | sum of squared distances between the given corners and the corners of the parametrised cuboid (FitCuboidBFGS.hs:51-66)
errfun :: Cuboid -> [Double] -> Double
errfun ps params = unsafePerformIO $ withCorners ps $ \pc -> withDoubles params $ \pp -> realToFrac <$> c_errfun pc pp

-- | [cx,cy,cz, a,b,c, q0,q1,q2,q3] -> 8 corners (FitCuboidBFGS.hs:98-114)
cuboidFromParams :: [Double] -> Cuboid
cuboidFromParams params = unsafePerformIO $ withDoubles params $ \pp -> allocaArray 24 $ \pc -> do
  _ <- c_cuboid_from_params pp pc
  toVecs <$> peekDoubles 24 pc
  where toVecs (x:y:z:r) = Vec3 x y z : toVecs r
        toVecs _         = []

fitWith :: Int -> Cuboid -> ([Double], Int, Double, Matrix Double)
fitWith variant ps = unsafePerformIO $ withCorners ps $ \pc ->
  allocaArray 10 $ \pout -> alloca $ \psteps -> alloca $ \perr -> allocaArray (cap * 13) $ \ppath -> do
    rc <- c_fit_cuboid pc (fromIntegral variant) pout psteps perr ppath (fromIntegral cap)
    if rc /= 0 then error "FitCuboidBFGS: hs_fit_cuboid failed" else do
      params <- peekDoubles 10 pout
      steps  <- fromIntegral <$> peek psteps
      err    <- realToFrac <$> (peek perr :: IO CDouble)
      let cols = if variant == 1 then 10 else 13
      path   <- peekDoubles (cap * cols) ppath
      return (params, steps, err, fromLists (takeWhile ((/= 0) . head) (rows cols path)))
  where cap = 4001   -- 2 stages x maxIt 2000 + 1 (FitCuboidBFGS.hs:184,201,233)
        rows n xs = case splitAt n xs of { (a, []) -> [a | not (null a)]; (a, b) -> a : rows n b }

-- | minimize NMSimplex2 1e-8 2000 (FitCuboidBFGS.hs:205-233 / :172-184 / :188-201); the optimiser is the library's NM-simplex2
fitCuboid, fitCuboidFromCenter, fitCuboidFromCenterFirst :: Cuboid -> ([Double], Int, Double, Matrix Double)
fitCuboid                = fitWith 0
fitCuboidFromCenter      = fitWith 1
fitCuboidFromCenterFirst = fitWith 2

fitCuboidFromCenterFirstError :: [Vec3] -> (Double, Int)
fitCuboidFromCenterFirstError ps = let (_, steps, err, _) = fitCuboidFromCenterFirst ps in (err, steps)

-- | the reference's self-check executable entry (FitCuboidBFGS.hs:255-282): fit the example box, print the result
main :: IO ()
main = do
  let box = cuboidFromParams [0, 0, 0, 1, 2, 3, 1, 0, 0, 0]
      (params, steps, err, _) = fitCuboid box
  print (params, steps, err)

-- ---------------------------------------------------------------------------------------------------------------- device path
-- | f = sum over the cloud of the squared distance to the nearest wall of the cuboid (planes of Main.hs:1852-1874, distance of
-- Main.hs:1371-1372), its gradient in the 10 parameters and the inliers per wall: one pass over the cloud on the GPU.
cuboidResidualGrad :: DeviceCloud -> [Double] -> IO (Double, [Double], [Int])
cuboidResidualGrad dc params = withCloudObjective dc ($ params)

-- | run an optimiser loop against ONE resident kernel: `withCloudObjective cloud $ \objective -> ... objective params ...`
withCloudObjective :: DeviceCloud -> (([Double] -> IO (Double, [Double], [Int])) -> IO a) -> IO a
withCloudObjective dc body = do
  n <- cloudSize dc
  bracket (beginEvalSession dc [0, n] False) endEvalSession $ \sess ->
    body $ \params -> do
      [rec] <- evalSessionEval sess [params]
      withDoubles params $ \pp -> withDoubles rec $ \pr -> alloca $ \pf -> allocaArray 10 $ \pg -> allocaArray 6 $ \pc -> do
        _ <- c_cuboid_grad_from_sums pp pr pf pg pc
        (,,) <$> (realToFrac <$> (peek pf :: IO CDouble)) <*> peekDoubles 10 pg <*> (map fromIntegral <$> peekArray 6 pc)

-- | BFGS over the whole room cloud (the library's optimiser driving a session): (params, iterations, f)
fitCuboidToCloudBFGS :: Ctx -> DeviceCloud -> [Double] -> IO ([Double], Int, Double)
fitCuboidToCloudBFGS ctx dc initial = withDoubles initial $ \pin -> allocaArray 10 $ \pout -> alloca $ \pf -> alloca $ \pit -> do
  check ctx =<< withCtxPtr ctx (\c -> withCloudPtr dc $ \pc -> c_fit_cuboid_cloud_bfgs c pc pin 200 1e-6 pout pf pit nullPtr)
  (,,) <$> peekDoubles 10 pout <*> (fromIntegral <$> peek pit) <*> (realToFrac <$> (peek pf :: IO CDouble))

-- | the reference's own optimiser over the whole room cloud: `minimize NMSimplex2 eps maxit` on the device objective; the library
-- posts the 11 corners of the initial simplex (and the 10 of every shrink step) to the session as one batch.
-- (params, iterations, f)
fitCuboidToCloudNM :: Ctx -> DeviceCloud -> [Double] -> [Double] -> Double -> Int -> IO ([Double], Int, Double)
fitCuboidToCloudNM ctx dc initial sizes eps maxIt =
  withDoubles initial $ \pin -> withDoubles sizes $ \pst -> allocaArray 10 $ \pout -> alloca $ \pf -> alloca $ \pit -> do
    check ctx =<< withCtxPtr ctx (\c -> withCloudPtr dc $ \pc ->
      c_fit_cuboid_cloud_nm c pc pin pst (realToFrac eps) (fromIntegral maxIt) pout pf pit nullPtr)
    (,,) <$> peekDoubles 10 pout <*> (fromIntegral <$> peek pit) <*> (realToFrac <$> (peek pf :: IO CDouble))
