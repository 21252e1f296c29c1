{-|
Module      : Crypto.Lol.Cyclotomic.Tensor.CUDA
Description : B200 back end for the 'Tensor' interface, next to Crypto.Lol.Cyclotomic.Tensor.CPP.

NOT COMPILED IN THIS REPOSITORY'S IMAGE (no GHC): written against lol-0.7.0.0, never type-checked.

Two layers:

 1. 'GT' -- the 'Tensor' instance.  libctensor_b200 exports the 29 symbols that
    Crypto.Lol.Cyclotomic.Tensor.CPP.Backend imports, with identical names and C signatures, so the whole of
    CPP.hs (marshalling, root tables, 'Dispatch') is reused: 'GT' is 'CT' built in a package that links
    @extra-libraries: ctensor_b200@ instead of compiling lol-cpp's @C-sources@ (see INTEGRATION.md for the cabal
    stanza).  Applications switch back ends the way they always have, by the proxy type
    (lol-cpp/examples/SHECPPMain.hs:21; rlwe-challenges/exec/RLWEChallengesMain.hs:30: @type T = GT@).

 2. Batch combinators -- one FFI crossing and one PCIe round trip for a whole list of ring elements, which is
    where the GPU pays off (a single 30 KiB element per call is dominated by launch and copy latency).
-}

{-# LANGUAGE DataKinds           #-}
{-# LANGUAGE FlexibleContexts    #-}
{-# LANGUAGE ScopedTypeVariables #-}

module Crypto.Lol.Cyclotomic.Tensor.CUDA
( GT
, crtBatch, crtInvBatch, lBatch, lInvBatch, mulGPowBatch, mulGDecBatch, mulBatch
) where

import Data.Int
import Data.Proxy
import qualified Data.Vector.Storable         as SV
import qualified Data.Vector.Storable.Mutable as SM
import System.IO.Unsafe (unsafePerformIO)

import Crypto.Lol.Cyclotomic.Tensor.CPP (CT)
import Crypto.Lol.Cyclotomic.Tensor.CUDA.Backend
import Crypto.Lol.Factored
import Crypto.Lol.Reflects
import Crypto.Lol.Types.Unsafe.ZqBasic (ZqBasic)

-- | The B200 tensor: representation and instance of 'CT', FFI resolved against libctensor_b200.
type GT = CT

-- | Run a chain of operators over many ring elements at once.  Elements are concatenated into one pinned
-- storable vector (Storable vectors are pinned, CPP.hs:329-336), transformed in place by the library and split again.
batchRq :: forall m q . (Fact m, Reflects q Int64)
        => String -> [SV.Vector (ZqBasic q Int64)] -> [SV.Vector (ZqBasic q Int64)]
batchRq ops xs = unsafePerformIO $ do
  let n     = proxy totientFact (Proxy :: Proxy m)
      pps   = [ (fromIntegral p, fromIntegral e) | (p, e) <- proxy ppsFact (Proxy :: Proxy m) ]
      q     = proxy value (Proxy :: Proxy q) :: Int64
      batch = length xs
  buf <- SV.thaw (SV.concat xs)
  withPlanRq pps [q] $ \plan ->
    SM.unsafeWith buf $ \p -> applyHostRq plan ops (castPtr' p) (fromIntegral batch)
  out <- SV.unsafeFreeze buf
  return [ SV.slice (i * n) n out | i <- [0 .. batch - 1] ]
  where castPtr' = Foreign.Ptr.castPtr

crtBatch, crtInvBatch, lBatch, lInvBatch, mulGPowBatch, mulGDecBatch
  :: forall m q . (Fact m, Reflects q Int64) => Proxy m -> [SV.Vector (ZqBasic q Int64)] -> [SV.Vector (ZqBasic q Int64)]
crtBatch     _ = batchRq "CRT"
crtInvBatch  _ = batchRq "CRTInv"
lBatch       _ = batchRq "L"
lInvBatch    _ = batchRq "LInv"
mulGPowBatch _ = batchRq "GPow"
mulGDecBatch _ = batchRq "GDec"

-- | Ring products of many element pairs through the CRT basis: CRT both operands, multiply coefficient-wise,
-- CRT^-1.  (Two batched calls plus a host zip; the fused device pipeline is SURVEY section 8(f) rank 1.)
mulBatch :: forall m q . (Fact m, Reflects q Int64, Num (ZqBasic q Int64))
         => Proxy m -> [SV.Vector (ZqBasic q Int64)] -> [SV.Vector (ZqBasic q Int64)] -> [SV.Vector (ZqBasic q Int64)]
mulBatch pm as bs = crtInvBatch pm $ zipWith (SV.zipWith (*)) (crtBatch pm as) (crtBatch pm bs)
