-- Drop-in for housescan/GroupConnectedComponents.hs: export list verbatim (GroupConnectedComponents.hs:1-3).  The type and the
-- output order are the reference's (components by ascending smallest vertex, edges inside a component in reverse input order:
-- GroupConnectedComponents.hs:46-54 `Map.fromListWith (++)`); the labelling runs on the GPU (union-find with atomics, canonical
-- min-index labels).  Not compiled in this repository's image (no GHC).
module GroupConnectedComponents
  ( groupConnectedComponents
  ) where

import Data.Int (Int32, Int64)
import Data.List (sortOn, groupBy)
import Data.Function (on)
import Data.Word (Word32)
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (allocaArray, peekArray, withArray)
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import Bijection (biject)
import HouseScanB200.Device
import HouseScanB200.FFI

-- | pure: the labels are canonical (smallest vertex index of the component), so the result does not depend on scheduling
groupConnectedComponents :: (Ord node) => [ ((node, node), a) ] -> [[ ((node, node), a) ]]
groupConnectedComponents [] = []
groupConnectedComponents edges = unsafePerformIO $
  withArray (map (fromIntegral . to . fst . fst) edges :: [Word32]) $ \ps ->
  withArray (map (fromIntegral . to . snd . fst) edges :: [Word32]) $ \pd ->
  allocaArray e $ \pcomp -> allocaArray e $ \porder -> alloca $ \pn -> do
    check defaultCtx =<< withCtxPtr defaultCtx (\c -> c_group_cc c ps pd (fromIntegral e) (fromIntegral nNodes) pcomp porder pn)
    comp  <- peekArray e pcomp  :: IO [Int32]
    order <- peekArray e porder :: IO [Int64]
    -- order lists the edge indices component by component in the reference's output order; comp !! i is the component of edge i
    let placed = [ (comp !! fromIntegral i, edges !! fromIntegral i) | i <- order ]
    return (map (map snd) (groupBy ((==) `on` fst) placed))
  where e          = length edges
        nodes      = concat [ [a, b] | ((a, b), _) <- edges ]
        (to, _)    = biject nodes                                           -- Bijection.hs:16-25: contiguous ids in Ord order
        nNodes     = length (groupBy (==) (sortOn id nodes))
