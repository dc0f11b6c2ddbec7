-- Managed handles over HouseScanB200.FFI: what the shim modules (FitCuboidBFGS, TranslationOptimizer, GroupConnectedComponents,
-- VectorUtil, HoniHelper) share.  Not compiled in this repository's image (no GHC); see INTEGRATION.md.
{-# LANGUAGE ScopedTypeVariables #-}
module HouseScanB200.Device
  ( Ctx, DeviceCloud, EvalSession
  , newCtx, defaultCtx, withCtxPtr, check, HsStatus(..)
  , uploadCloud, downloadCloud, cloudSize, withCloudPtr, allocCloud
  , withDoubles, peekDoubles
  , beginEvalSession, evalSessionEval, evalSessionPost, evalSessionWait, endEvalSession
  ) where

import Control.Exception (throwIO, ErrorCall(..))
import Control.Monad (when)
import Data.Int (Int32, Int64)
import qualified Data.Vector.Storable as V
import qualified Data.Vector.Storable.Mutable as VM
import Data.Vect.Float (Vec3(..))
import Foreign.C.String (peekCString)
import Foreign.C.Types (CDouble(..), CFloat(..))
import Foreign.ForeignPtr
import qualified Foreign.Concurrent as FC
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (withArrayLen, peekArray, allocaArray)
import Foreign.Ptr (Ptr, castPtr, nullPtr)
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import HouseScanB200.FFI

-- | hs_status (include/housescan_b200.h): every failure carries hs_last_error's text
data HsStatus = HsOk | HsEInval | HsECuda | HsENccl | HsENoMem | HsESingular | HsEIO deriving (Eq, Show, Enum)

newtype Ctx = Ctx (ForeignPtr HsCtx)
data DeviceCloud = DeviceCloud Ctx (ForeignPtr HsCloud)   -- `Vector Vec3` living in HBM (Main.hs:117-121 cloudPoints)
data EvalSession = EvalSession Ctx (Ptr HsEvalSession) Int -- resident kernel + number of rooms

-- | One CUDA device + stream.  Fails (HS_ECUDA) when there is no sm_100 GPU: there is no CPU fallback.
newCtx :: Int -> IO Ctx
newCtx dev = alloca $ \pp -> do
  rc <- c_ctx_create (fromIntegral dev) pp
  when (rc /= 0) $ do
    msg <- peekCString =<< c_last_error nullPtr
    throwIO (ErrorCall ("housescan_b200: " ++ msg))
  p <- peek pp
  Ctx <$> newForeignPtr_ p   -- destroyed explicitly at exit (hs_ctx_destroy synchronises the device; not a finaliser job)

-- | The context of the single GLUT thread (Main.hs:1143-1159); ghci users of `run` (Main.hs:1184) create their own.
defaultCtx :: Ctx
defaultCtx = unsafePerformIO (newCtx 0)
{-# NOINLINE defaultCtx #-}

withCtxPtr :: Ctx -> (Ptr HsCtx -> IO a) -> IO a
withCtxPtr (Ctx fp) = withForeignPtr fp

-- | status -> the reference's error behaviour (`error` with the library's message; INTEGRATION.md section 3)
check :: Ctx -> Int32 -> IO ()
check _ 0 = return ()
check ctx rc = withCtxPtr ctx $ \p -> do
  msg <- peekCString =<< c_last_error p
  throwIO (ErrorCall (show (toEnum (fromIntegral rc) :: HsStatus) ++ ": " ++ msg))

-- | `Vector Vec3` is 12 bytes per point AoS (Main.hs:39-42 Storable Vec3): the buffer goes to the device as it lies in memory
uploadCloud :: Ctx -> V.Vector Vec3 -> IO DeviceCloud
uploadCloud ctx v = V.unsafeWith v $ \pv -> alloca $ \pp -> do
  check ctx =<< withCtxPtr ctx (\c -> c_cloud_upload c (castPtr pv) (fromIntegral (V.length v)) pp)
  wrapCloud ctx =<< peek pp

allocCloud :: Ctx -> Int -> IO DeviceCloud
allocCloud ctx n = alloca $ \pp -> do
  check ctx =<< withCtxPtr ctx (\c -> c_cloud_alloc c (fromIntegral n) pp)
  wrapCloud ctx =<< peek pp

wrapCloud :: Ctx -> Ptr HsCloud -> IO DeviceCloud
wrapCloud ctx@(Ctx cfp) p = do
  fp <- FC.newForeignPtr p (withForeignPtr cfp $ \c -> c_cloud_free c p >> return ())   -- hs_cloud_free when the last reference dies
  return (DeviceCloud ctx fp)

downloadCloud :: DeviceCloud -> IO (V.Vector Vec3)
downloadCloud dc@(DeviceCloud ctx _) = do
  n <- cloudSize dc
  out <- VM.new n
  check ctx =<< withCtxPtr ctx (\c -> withCloudPtr dc $ \pc -> VM.unsafeWith out $ \po -> c_cloud_download c pc (castPtr po))
  V.unsafeFreeze out

cloudSize :: DeviceCloud -> IO Int
cloudSize dc = fromIntegral <$> withCloudPtr dc c_cloud_size

withCloudPtr :: DeviceCloud -> (Ptr HsCloud -> IO a) -> IO a
withCloudPtr (DeviceCloud _ fp) = withForeignPtr fp

withDoubles :: [Double] -> (Ptr CDouble -> IO a) -> IO a
withDoubles xs k = withArrayLen (map realToFrac xs) (\_ p -> k p)

peekDoubles :: Int -> Ptr CDouble -> IO [Double]
peekDoubles n p = map realToFrac <$> peekArray n p

-- | Evaluation sessions (include/housescan_b200.h): the optimiser loop of FitCuboidBFGS.hs:184,201,233 evaluates its objective up
-- to 2000 times per stage; a session keeps ONE kernel resident and turns an evaluation into a parameter post + a record read.
beginEvalSession :: DeviceCloud -> [Int] -> Bool -> IO EvalSession
beginEvalSession dc@(DeviceCloud ctx _) roomOffsets allreduce =
  withArrayLen (map fromIntegral roomOffsets :: [Int64]) $ \len po -> alloca $ \pp -> do
    check ctx =<< withCtxPtr ctx (\c -> withCloudPtr dc $ \pc ->
      c_eval_session_begin c pc po (fromIntegral (len - 1)) (if allreduce then 1 else 0) pp)
    s <- peek pp
    return (EvalSession ctx s (len - 1))

-- | one evaluation: params = nrooms x 10 doubles -> nrooms x 24-double records
evalSessionEval :: EvalSession -> [[Double]] -> IO [[Double]]
evalSessionEval (EvalSession ctx s nr) params = withDoubles (concat params) $ \pp -> allocaArray (nr * 24) $ \pr -> do
  check ctx =<< c_eval_session_eval s pp pr
  chunk 24 <$> peekDoubles (nr * 24) pr

-- | post `count` evaluations ahead (simplex vertices, finite differences, line-search points); returns nothing, see evalSessionWait
evalSessionPost :: EvalSession -> [[[Double]]] -> IO ()
evalSessionPost (EvalSession ctx s _) batch = withDoubles (concatMap concat batch) $ \pp ->
  check ctx =<< c_eval_session_post s pp (fromIntegral (length batch))

evalSessionWait :: EvalSession -> Int -> IO [[Double]]
evalSessionWait (EvalSession ctx s nr) seqNo = allocaArray (nr * 24) $ \pr -> do
  check ctx =<< c_eval_session_wait s (fromIntegral seqNo) pr
  chunk 24 <$> peekDoubles (nr * 24) pr

endEvalSession :: EvalSession -> IO ()
endEvalSession (EvalSession ctx s _) = check ctx =<< c_eval_session_end s

chunk :: Int -> [a] -> [[a]]
chunk _ [] = []
chunk n xs = let (a, b) = splitAt n xs in a : chunk n b
