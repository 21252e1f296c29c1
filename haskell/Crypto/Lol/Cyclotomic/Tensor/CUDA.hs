{-|
Module      : Crypto.Lol.Cyclotomic.Tensor.CUDA
Description : B200 back end for the 'Tensor' interface, next to Crypto.Lol.Cyclotomic.Tensor.CPP.

NOT COMPILED IN THIS REPOSITORY'S IMAGE (no GHC / stack / cabal, no network): written against lol-0.7.0.0 and
lol-cpp-0.0.0.4 method by method from lol/Crypto/Lol/Cyclotomic/Tensor.hs:86-193 and
lol-cpp/Crypto/Lol/Cyclotomic/Tensor/CPP.hs:204-264, never type-checked.  A maintainer with GHC should expect to fix
imports and constraint plumbing; the C side of every call below is exercised through the same symbols by the Python
tests of this repository (lol_b200/capi.py, tests/test_gpu_*.py).

What differs from 'CT' (CPP.hs:93-95): a 'GT' holds its coefficients in GPU MEMORY.

@
  data GT m r = GT (GT' m r)          -- device-resident: ForeignPtr to [totient m] elements of r on the GPU,
              | GZ (IZipVector m r)   --   freed by lolb_dev_free; the layout is CT's (Backend.hs:80-90)
@

Every 'Tensor' method that the reference implements in C (l, lInv, mulG*, divG*, crt, crtInv, mulGCRT, divGCRT,
tGaussianDec, gSqNormDec) and every method it implements in Haskell over index vectors (twacePowDec, embedPow,
embedDec, twaceCRT, embedCRT, coeffs, powBasisPow) is ONE kernel launch on the device pointer
(include/lol_b200.h, `lolb_*`), so a chain such as @crtInv . mulGCRT . crt@ crosses PCIe only when the caller finally
reads the coefficients ('Foldable' / 'Show' / 'Eq' / protobuf / 'fmapT' with an arbitrary closure).  Those readers
download ('gtToVector'); 'fmapT' / 'zipWithT' re-upload.  The coefficient-wise maps Lol actually uses between
transforms -- lift, reduce, rescale, roundCoset -- have device versions in the library (lolb_liftRq ...); they are
exported here as 'liftT', 'reduceT', 'rescaleDropT' for a `Cyc`-level specialisation to pick up (RULES / class
methods: a design decision for the maintainer).

Plans (root tables, twiddles, gCRT vectors on the device) are created once per (m, moduli) and memoised, like
CPP.hs memoises 'ru' / 'ruInv' (CPP.hs:422-442).  The library derives omega exactly as ZqBasic does
(ZqBasic.hs:144-165: smallest generator of Z_q^*), so results equal 'CT' bit for bit over Z_q.
-}

{-# LANGUAGE ConstraintKinds           #-}
{-# LANGUAGE DataKinds                 #-}
{-# LANGUAGE FlexibleContexts          #-}
{-# LANGUAGE FlexibleInstances         #-}
{-# LANGUAGE GADTs                     #-}
{-# LANGUAGE InstanceSigs              #-}
{-# LANGUAGE KindSignatures            #-}
{-# LANGUAGE MultiParamTypeClasses     #-}
{-# LANGUAGE PolyKinds                 #-}
{-# LANGUAGE RankNTypes                #-}
{-# LANGUAGE RoleAnnotations           #-}
{-# LANGUAGE ScopedTypeVariables       #-}
{-# LANGUAGE StandaloneDeriving        #-}
{-# LANGUAGE TypeFamilies              #-}
{-# LANGUAGE TypeOperators             #-}
{-# LANGUAGE UndecidableInstances      #-}

module Crypto.Lol.Cyclotomic.Tensor.CUDA
( GT
  -- * device versions of the coefficient-wise maps Lol applies between transforms
, liftT, reduceT, rescaleDropT
  -- * whole batches through one host call (Storable vectors in, Storable vectors out)
, crtBatch, crtInvBatch
) where

import Control.Applicative    hiding ((*>))
import Control.DeepSeq
import Control.Monad.Random
import Data.Coerce
import Data.Constraint        hiding ((***))
import Data.Foldable          as F
import Data.Int
import Data.IORef
import qualified Data.Map.Strict as M
import Data.Maybe
import Data.Traversable       as T
import qualified Data.Vector.Storable         as SV
import qualified Data.Vector.Storable.Mutable as SM
import Foreign.ForeignPtr
import Foreign.Ptr
import Foreign.Storable       (Storable, sizeOf)
import System.IO.Unsafe       (unsafePerformIO)

import Crypto.Lol.CRTrans
import Crypto.Lol.Cyclotomic.Tensor
import Crypto.Lol.Cyclotomic.Tensor.CPP             (CT)         -- host fallback for crtSetDec / Module (GF fp d)
import Crypto.Lol.Cyclotomic.Tensor.CPP.Instances   ()           -- orphan Storable (a,b), Complex (Backend.hs:80-90)
import Crypto.Lol.Cyclotomic.Tensor.CUDA.Backend
import Crypto.Lol.Prelude                           as LP
import Crypto.Lol.Reflects
import Crypto.Lol.Types.FiniteField
import Crypto.Lol.Types.IZipVector
import Crypto.Lol.Types.Proto
import Crypto.Lol.Types.Unsafe.Complex
import Crypto.Lol.Types.Unsafe.RRq
import Crypto.Lol.Types.Unsafe.ZqBasic

-- ------------------------------------------------------------------------------------------------ representation

-- | @totient m@ coefficients of type @r@ in GPU memory, in the element layout of CPP/Backend.hs:80-90 (RNS tuples
-- interleaved by limb).  The finalizer is @lolb_dev_free@.
newtype GT' (m :: Factored) r = GT' (ForeignPtr r)
type role GT' representational nominal

-- | An implementation of 'Tensor' backed by libctensor_b200 (CUDA, sm_100a).  Same two constructors as 'CT'
-- (CPP.hs:93-95): a flat device array for element types the library serves, a boxed vector for everything else.
data GT (m :: Factored) r where
  GT :: DevElt r => GT' m r -> GT m r
  GZ :: IZipVector m r -> GT m r

-- | Element types the device path serves, and which C symbols serve them: the analogue of `Dispatch`
-- (CPP/Backend.hs:92-302).  @RNS@ instances recurse over pairs exactly as `Tuple` does there.
class Storable r => DevElt r where
  -- | plan for index @m@ over this element type (memoised; Rq plans carry the moduli, the others are modulus-free)
  planFor  :: Fact m => proxy m -> Tagged r Plan
  -- | LOLB_RING_* tag of the ring-extension operators
  ringTag  :: Tagged r Int32
  -- | symbol suffix selector for the single-index operators
  opL, opLInv, opGPow, opGDec :: Tagged r DevOp
  opGInvPow, opGInvDec        :: Tagged r DevOpStatus
  opCRT, opCRTInv             :: Tagged r (Maybe DevOp)        -- Nothing: no CRT over this type (Int64, Double)
  opMul                       :: Tagged r (Maybe DevOp2)
  opNormSq                    :: Tagged r (Maybe DevOpOut)

instance (Reflects q Int64) => DevElt (ZqBasic q Int64) where
  planFor pm = tag $ memoPlanRq (ppsOf pm) [proxy value (Proxy :: Proxy q)]
  ringTag    = tag ringRq
  opL = tag c_lRq; opLInv = tag c_lInvRq; opGPow = tag c_gPowRq; opGDec = tag c_gDecRq
  opGInvPow = tag c_gInvPowRq; opGInvDec = tag c_gInvDecRq
  opCRT = tag (Just c_crtRq); opCRTInv = tag (Just c_crtInvRq)
  opMul = tag (Just c_mulRq)
  opNormSq = tag Nothing

-- RNS pairs: one plan over the concatenated modulus list, tupSize = number of limbs (Backend.hs:104-121 `Tuple`)
instance (DevElt a, DevElt b, Moduli a, Moduli b) => DevElt (a, b) where
  planFor pm = tag $ memoPlanRq (ppsOf pm) (proxy moduli (Proxy :: Proxy (a, b)))
  ringTag    = tag ringRq
  opL = tag c_lRq; opLInv = tag c_lInvRq; opGPow = tag c_gPowRq; opGDec = tag c_gDecRq
  opGInvPow = tag c_gInvPowRq; opGInvDec = tag c_gInvDecRq
  opCRT = tag (Just c_crtRq); opCRTInv = tag (Just c_crtInvRq)
  opMul = tag (Just c_mulRq)
  opNormSq = tag Nothing

instance DevElt Int64 where
  planFor pm = tag $ memoPlanC (ppsOf pm) 1
  ringTag    = tag ringInt
  opL = tag c_lR; opLInv = tag c_lInvR; opGPow = tag c_gPowR; opGDec = tag c_gDecR
  opGInvPow = tag c_gInvPowR; opGInvDec = tag c_gInvDecR
  opCRT = tag Nothing; opCRTInv = tag Nothing; opMul = tag Nothing
  opNormSq = tag (Just c_normSqR)

instance DevElt Double where
  planFor pm = tag $ memoPlanC (ppsOf pm) 1
  ringTag    = tag ringDouble
  opL = tag c_lDouble; opLInv = tag c_lInvDouble
  opGPow = tag noOp; opGDec = tag noOp; opGInvPow = tag noOpS; opGInvDec = tag noOpS      -- Backend.hs:267-283: not dispatched
  opCRT = tag Nothing; opCRTInv = tag Nothing; opMul = tag Nothing
  opNormSq = tag (Just c_normSqD)

instance DevElt (Complex Double) where
  planFor pm = tag $ memoPlanC (ppsOf pm) 1
  ringTag    = tag ringComplex
  opL = tag c_lC; opLInv = tag c_lInvC; opGPow = tag c_gPowC; opGDec = tag c_gDecC
  opGInvPow = tag c_gInvPowC; opGInvDec = tag c_gInvDecC
  opCRT = tag (Just c_crtC); opCRTInv = tag (Just c_crtInvC)
  opMul = tag (Just c_mulC)
  opNormSq = tag Nothing

-- | The modulus list of an RNS tuple, leftmost limb first (the order `Tuple` marshals them, Backend.hs:104-121).
class Moduli r where moduli :: Tagged r [Int64]
instance Reflects q Int64 => Moduli (ZqBasic q Int64) where moduli = tag [proxy value (Proxy :: Proxy q)]
instance (Moduli a, Moduli b) => Moduli (a, b) where
  moduli = tag $ proxy moduli (Proxy :: Proxy a) ++ proxy moduli (Proxy :: Proxy b)

ppsOf :: forall m proxy . Fact m => proxy m -> [CPP]
ppsOf _ = [ (fromIntegral p, fromIntegral e) | (p, e) <- proxy ppsFact (Proxy :: Proxy m) ]

-- ------------------------------------------------------------------------------------------------ plans, memoised

{-# NOINLINE planCache #-}
planCache :: IORef (M.Map (Bool, [CPP], [Int64], Int) Plan)
planCache = unsafePerformIO $ newIORef M.empty

memoPlan :: (Bool, [CPP], [Int64], Int) -> IO Plan -> Plan
memoPlan key mk = unsafePerformIO $ do
  cache <- readIORef planCache
  case M.lookup key cache of
    Just p  -> return p
    Nothing -> do p <- mk
                  atomicModifyIORef' planCache (\c -> (M.insert key p c, ()))
                  return p

memoPlanRq :: [CPP] -> [Int64] -> Plan
memoPlanRq pps qs = memoPlan (True, pps, qs, length qs) (newPlanRq pps qs)

memoPlanC :: [CPP] -> Int -> Plan
memoPlanC pps k = memoPlan (False, pps, [], k) (newPlanC pps k)

-- ------------------------------------------------------------------------------------------------ host <-> device

totM :: forall m proxy . Fact m => proxy m -> Int
totM _ = proxy totientFact (Proxy :: Proxy m)

newDev :: forall r . Storable r => Int -> IO (ForeignPtr r)
newDev n = do p <- devAlloc (fromIntegral (n * sizeOf (undefined :: r)))
              newForeignPtr p_devFree (castPtr p)

-- | upload a host vector (one 'SV.thaw'-free copy: Storable vectors are contiguous, Backend.hs:68-74)
toDev :: forall m r . Storable r => SV.Vector r -> GT' m r
toDev v = unsafePerformIO $ do
  fp <- newDev (SV.length v)
  SV.unsafeWith v $ \src -> withForeignPtr fp $ \dst ->
    devUpload (castPtr dst) (castPtr src) (fromIntegral (SV.length v * sizeOf (undefined :: r)))
  return (GT' fp)

-- | download: the only place a device tensor crosses PCIe
gtToVector :: forall m r . (Fact m, Storable r) => GT' m r -> SV.Vector r
gtToVector (GT' fp) = unsafePerformIO $ do
  let n = totM (Proxy :: Proxy m)
  buf <- SM.new n
  SM.unsafeWith buf $ \dst -> withForeignPtr fp $ \src ->
    devDownload (castPtr dst) (castPtr src) (fromIntegral (n * sizeOf (undefined :: r)))
  SV.unsafeFreeze buf

toGT :: (Fact m, DevElt r) => GT m r -> GT m r
toGT v@(GT _) = v
toGT (GZ v)   = GT $ toDev $ SV.convert $ unIZipVector v

toGZ :: Fact m => GT m r -> GT m r
toGZ (GT v)   = GZ $ fromMaybe (error "toGZ: internal error") $ iZipVector $ SV.convert $ gtToVector v
toGZ v@(GZ _) = v

-- | a pure device operator: fresh output buffer (device-to-device copy), one kernel on it, nothing retained
unaryDev :: forall m r . (Fact m, DevElt r) => DevOp -> GT' m r -> GT' m r
unaryDev op (GT' src) = unsafePerformIO $ do
  let n = totM (Proxy :: Proxy m)
  dst <- newDev n
  withForeignPtr src $ \ps -> withForeignPtr dst $ \pd -> do
    devCopy (castPtr pd) (castPtr ps) (fromIntegral (n * sizeOf (undefined :: r)))
    withPlan (proxy (planFor (Proxy :: Proxy m)) (Proxy :: Proxy r)) $ \pl -> op pl (castPtr pd) 1 nullPtr >>= check "lolb_tensor*"
  return (GT' dst)

wrap1 :: (Fact m, DevElt r) => (GT' m r -> GT' m r) -> GT m r -> GT m r
wrap1 f v = case toGT v of GT x -> GT (f x); _ -> error "wrap1: unreachable"

-- ------------------------------------------------------------------------------------------------ class instances

instance (Fact m, Eq r, Storable r) => Eq (GT m r) where
  a == b = unwrapHost a == unwrapHost b

unwrapHost :: (Fact m) => GT m r -> IZipVector m r
unwrapHost v = case toGZ v of GZ z -> z; _ -> error "unwrapHost: unreachable"

instance (Fact m, Show r) => Show (GT m r) where show = show . unwrapHost
instance (Fact m, NFData r) => NFData (GT m r) where
  rnf (GT (GT' fp)) = fp `seq` ()          -- device buffers are strict by construction
  rnf (GZ v)        = rnf v

instance (Protoable (IZipVector m r), Fact m, DevElt r) => Protoable (GT m r) where
  type ProtoType (GT m r) = ProtoType (IZipVector m r)           -- wire format of CPP.hs:115-121, lol/Lol.proto:20-49
  toProto   = toProto . unwrapHost
  fromProto x = toGT . GZ <$> fromProto x

instance Fact m => Functor (GT m) where fmap f x = pure f <*> x
instance Fact m => Applicative (GT m) where
  pure = GZ . pure
  f <*> a = GZ (unwrapHost f <*> unwrapHost a)
instance Fact m => Foldable (GT m) where foldMap = foldMapDefault
instance Fact m => Traversable (GT m) where traverse f v = GZ <$> T.traverse f (unwrapHost v)

instance (Additive r, DevElt r, Fact m) => Additive.C (GT m r) where
  a + b    = zipWithT (+) a b
  negate   = fmapT negate
  zero     = GZ (pure zero)
instance (ZeroTestable r, Fact m) => ZeroTestable.C (GT m r) where isZero = isZero . unwrapHost
instance (Random r, DevElt r, Fact m) => Random (GT m r) where
  random = runRand $ toGT . GZ <$> liftRand random
  randomR = error "randomR nonsensical for GT"
instance (GFCtx fp d, Fact m, Additive (GT m fp), DevElt fp) => Module.C (GF fp d) (GT m fp) where
  r *> v = toGT $ GZ $ fromJust $ iZipVector $ LP.fromList' $ unCoeffs $ r *> Coeffs (F.toList (unwrapHost v))

-- ------------------------------------------------------------------------------------------------ instance Tensor GT

instance Tensor GT where

  type TElt GT r = DevElt r

  entailIndexT  = tag $ Sub Dict
  entailEqT     = tag $ Sub Dict
  entailZTT     = tag $ Sub Dict
  entailNFDataT = tag $ Sub Dict
  entailRandomT = tag $ Sub Dict
  entailShowT   = tag $ Sub Dict
  entailModuleT = tag $ Sub Dict

  -- Tensor.hs:112: the scalar in the powerful basis = [r, 0, 0, ...]; built on the host (n words), uploaded once
  scalarPow :: forall m r . (Additive r, Fact m, DevElt r) => r -> GT m r
  scalarPow r = GT $ toDev $ SV.generate (totM (Proxy :: Proxy m)) (\i -> if i == 0 then r else zero)

  -- l.cpp:109-180 -> lolb_tensorL* / lolb_tensorLInv* (k_line_stream / fused_plain)
  l    = wrap1 $ \x -> unaryDev (proxy opL    (elt x)) x
  lInv = wrap1 $ \x -> unaryDev (proxy opLInv (elt x)) x

  -- g.cpp:125-155
  mulGPow = wrap1 $ \x -> unaryDev (proxy opGPow (elt x)) x
  mulGDec = wrap1 $ \x -> unaryDev (proxy opGDec (elt x)) x

  -- g.cpp:169-273: status 0 -> Nothing (CPP.hs:321-323)
  divGPow v = case toGT v of GT x -> GT <$> statusDev (proxy opGInvPow (elt x)) x
  divGDec v = case toGT v of GT x -> GT <$> statusDev (proxy opGInvDec (elt x)) x

  -- Tensor.hs:131-140; CPP.hs:225-231.  `crtInfo` decides existence exactly as for CT (CRTrans mon r); the root
  -- tables themselves live in the plan.  mulGCRT / divGCRT are lolb_mulRq against the plan's device-resident
  -- gCRT / gInvCRT vectors (Tensor.hs:319-337) -- no operand upload.
  crtFuncs :: forall mon m r . (CRTrans mon r, Fact m, DevElt r)
           => mon (r -> GT m r, GT m r -> GT m r, GT m r -> GT m r, GT m r -> GT m r, GT m r -> GT m r)
  crtFuncs = do
    (_ :: CRTInfo r) <- proxyT crtInfo (Proxy :: Proxy m)          -- Nothing / error exactly where CT has no CRT
    let pr       = Proxy :: Proxy r
        Just fwd = proxy opCRT pr
        Just inv = proxy opCRTInv pr
    return ( \r -> GT $ toDev $ SV.replicate (totM (Proxy :: Proxy m)) r        -- scalarCRT (CPP.hs:226 `repl`)
           , wrap1 (mulByPlanVector False)                                      -- mulGCRT
           , wrap1 (mulByPlanVector True)                                       -- divGCRT
           , wrap1 (unaryDev fwd)                                               -- crt     crt.cpp:562-566
           , wrap1 (unaryDev inv) )                                             -- crtInv  crt.cpp:569-581

  -- CPP.hs:376-389 + GaussRandom.hs:34-59: the draw happens ON THE DEVICE (lolb_tGaussianDec: Philox4x32-10 + polar
  -- Box-Muller), seeded from the caller's MonadRandom, so the sample never exists in host memory
  tGaussianDec :: forall v rnd m q . (OrdFloat q, Random q, DevElt q, ToRational v, Fact m, MonadRandom rnd)
               => v -> rnd (GT m q)
  tGaussianDec v = do
    seed <- getRandom
    return $ GT $ gaussianDev (Proxy :: Proxy m) (realToField v :: Double) (seed :: Word64)

  -- norm.cpp:39-80 -> lolb_tensorNormSqR / D (one scalar comes back)
  gSqNormDec v = case toGT v of GT x -> normSqDev x

  -- Extension.hs:54-129 (Haskell over index vectors in the reference) -> one gather kernel each (ext_stream.cu)
  twacePowDec = wrapExt c_twacePowDec
  embedPow    = wrapExt c_embedPow
  embedDec    = wrapExt c_embedDec

  crtExtFuncs :: forall mon m m' r . (CRTrans mon r, m `Divides` m', DevElt r)
              => mon (GT m' r -> GT m r, GT m r -> GT m' r)
  crtExtFuncs = do
    (_ :: CRTInfo r) <- proxyT crtInfo (Proxy :: Proxy m')
    return (wrapExt c_twaceCRT, wrapExt c_embedCRT)

  -- Extension.hs:90-93: phi'/phi elements of O_m, produced back to back by one kernel, sliced on the device
  coeffs v = case toGT v of GT x -> GT <$> coeffsDev x

  -- Extension.hs:133-143: lolb_powBasisPow
  powBasisPow :: forall m m' r . (Ring r, DevElt r, m `Divides` m') => Tagged m [GT m' r]
  powBasisPow = tag $ GT <$> powBasisDev (Proxy :: Proxy '(m, m'))

  -- Extension.hs:145-164 is finite-field arithmetic on the host in the reference as well: computed by CT, uploaded
  crtSetDec :: forall m m' fp . (m `Divides` m', PrimeField fp, Coprime (PToF (CharOf fp)) m', DevElt fp)
            => Tagged m [GT m' fp]
  crtSetDec = tag $ (GT . toDev . ctToVector) <$> proxy (crtSetDec :: Tagged m [CT m' fp]) (Proxy :: Proxy m)

  -- arbitrary host closures cannot run on the GPU: download, map, upload.  Lol's own closures (lift, reduce,
  -- rescale, roundCoset) have device versions below.
  fmapT f v = toGT $ GZ $ fmap f (unwrapHost v)
  zipWithT f a b = toGT $ GZ $ f <$> unwrapHost a <*> unwrapHost b
  unzipT v = let z = unwrapHost v in (toGT $ GZ $ fst <$> z, toGT $ GZ $ snd <$> z)

elt :: GT' m r -> Proxy r
elt _ = Proxy

-- ------------------------------------------------------------------------------------------------ helpers over the FFI

statusDev :: forall m r . (Fact m, DevElt r) => DevOpStatus -> GT' m r -> Maybe (GT' m r)
statusDev op (GT' src) = unsafePerformIO $ do
  let n = totM (Proxy :: Proxy m)
  dst <- newDev n
  ok <- withForeignPtr src $ \ps -> withForeignPtr dst $ \pd -> do
    devCopy (castPtr pd) (castPtr ps) (fromIntegral (n * sizeOf (undefined :: r)))
    withPlan (proxy (planFor (Proxy :: Proxy m)) (Proxy :: Proxy r)) $ \pl -> op pl (castPtr pd)
  return $ if ok then Just (GT' dst) else Nothing

mulByPlanVector :: forall m r . (Fact m, DevElt r) => Bool -> GT' m r -> GT' m r
mulByPlanVector inverse (GT' src) = unsafePerformIO $ do
  let n = totM (Proxy :: Proxy m)
  dst <- newDev n
  withForeignPtr src $ \ps -> withForeignPtr dst $ \pd -> do
    devCopy (castPtr pd) (castPtr ps) (fromIntegral (n * sizeOf (undefined :: r)))
    withPlan (proxy (planFor (Proxy :: Proxy m)) (Proxy :: Proxy r)) $ \pl -> do
      g <- c_planGcrtDev pl (if inverse then 1 else 0)
      c_mulRq pl (castPtr pd) g 1 1 nullPtr >>= check "lolb_mulRq"
  return (GT' dst)

gaussianDev :: forall m q . (Fact m, DevElt q) => Proxy m -> Double -> Word64 -> GT' m q
gaussianDev pm v seed = unsafePerformIO $ do
  dst <- newDev (totM pm)
  withForeignPtr dst $ \pd -> withPlan (proxy (planFor pm) (Proxy :: Proxy q)) $ \pl ->
    c_tGaussianDec pl v seed 0 (castPtr pd) 1 nullPtr >>= check "lolb_tGaussianDec"
  return (GT' dst)

normSqDev :: forall m r . (Fact m, DevElt r) => GT' m r -> r
normSqDev (GT' src) = unsafePerformIO $ do
  let Just op = proxy opNormSq (Proxy :: Proxy r)
  out <- newDev 1
  withForeignPtr src $ \ps -> withForeignPtr out $ \po ->
    withPlan (proxy (planFor (Proxy :: Proxy m)) (Proxy :: Proxy r)) $ \pl -> op pl (castPtr ps) (castPtr po) 1 nullPtr >>= check "lolb_tensorNormSq*"
  SV.head <$> (return $! gtToVector1 out)
  where gtToVector1 fp = unsafePerformIO $ do
          buf <- SM.new 1
          SM.unsafeWith buf $ \d -> withForeignPtr fp $ \s -> devDownload (castPtr d) (castPtr s) (fromIntegral (sizeOf (undefined :: r)))
          SV.unsafeFreeze buf

-- | the two-index operators: one memoised @lolb_ext@ per '(m, m', element type)
wrapExt :: forall a b r . (Fact a, Fact b, DevElt r) => ExtOp -> GT a r -> GT b r
wrapExt op v = case toGT v of
  GT (GT' src) -> GT $ unsafePerformIO $ do
    dst <- newDev (totM (Proxy :: Proxy b))
    withForeignPtr src $ \ps -> withForeignPtr dst $ \pd ->
      withExt (proxy (planFor (Proxy :: Proxy a)) (Proxy :: Proxy r)) (proxy (planFor (Proxy :: Proxy b)) (Proxy :: Proxy r)) $ \ext ->
        op ext (proxy ringTag (Proxy :: Proxy r)) (castPtr ps) (castPtr pd) 1 nullPtr >>= check "lolb_ext operator"
    return (GT' dst)
  _ -> error "wrapExt: unreachable"

coeffsDev :: forall m m' r . (m `Divides` m', DevElt r) => GT' m' r -> [GT' m r]
coeffsDev (GT' src) = unsafePerformIO $ do
  let n = totM (Proxy :: Proxy m); n' = totM (Proxy :: Proxy m'); sz = sizeOf (undefined :: r)
  flat <- newDev n'
  withForeignPtr src $ \ps -> withForeignPtr flat $ \pf ->
    withExt (proxy (planFor (Proxy :: Proxy m)) (Proxy :: Proxy r)) (proxy (planFor (Proxy :: Proxy m')) (Proxy :: Proxy r)) $ \ext ->
      c_coeffsPowDec ext (proxy ringTag (Proxy :: Proxy r)) (castPtr ps) (castPtr pf) 1 nullPtr >>= check "lolb_coeffsPowDec"
  forM [0 .. n' `div` n - 1] $ \i -> do
    part <- newDev n
    withForeignPtr flat $ \pf -> withForeignPtr part $ \pp ->
      devCopy (castPtr pp) (castPtr pf `plusPtr` (i * n * sz)) (fromIntegral (n * sz))
    return (GT' part)

powBasisDev :: forall m m' r . (m `Divides` m', DevElt r) => Proxy '(m, m') -> [GT' m' r]
powBasisDev _ = unsafePerformIO $ do
  let n = totM (Proxy :: Proxy m); n' = totM (Proxy :: Proxy m'); sz = sizeOf (undefined :: r); rel = n' `div` n
  flat <- newDev (rel * n')
  withForeignPtr flat $ \pf ->
    withExt (proxy (planFor (Proxy :: Proxy m)) (Proxy :: Proxy r)) (proxy (planFor (Proxy :: Proxy m')) (Proxy :: Proxy r)) $ \ext ->
      c_powBasisPow ext (proxy ringTag (Proxy :: Proxy r)) (castPtr pf) nullPtr >>= check "lolb_powBasisPow"
  forM [0 .. rel - 1] $ \i -> do
    part <- newDev n'
    withForeignPtr flat $ \pf -> withForeignPtr part $ \pp ->
      devCopy (castPtr pp) (castPtr pf `plusPtr` (i * n' * sz)) (fromIntegral (n' * sz))
    return (GT' part)

ctToVector :: (Fact m, Storable r) => CT m r -> SV.Vector r
ctToVector = SV.fromList . F.toList

-- ------------------------------------------------------------------------------------------------ device fmapT closures

-- | @fmapT lift@ (UCyc.hs:267-283), @fmapT reduce@ (UCyc.hs:285-296) and the RNS limb drop of @rescaleCyc@
-- (Cyc.hs:529-541) without leaving the device: lolb_liftRq / lolb_reduceRq / lolb_rescaleDropRq.
liftT :: forall m q . (Fact m, Reflects q Int64) => GT m (ZqBasic q Int64) -> GT m Int64
liftT v = case toGT v of
  GT (GT' src) -> GT $ unsafePerformIO $ do
    dst <- newDev (totM (Proxy :: Proxy m))
    withForeignPtr src $ \ps -> withForeignPtr dst $ \pd ->
      withPlan (proxy (planFor (Proxy :: Proxy m)) (Proxy :: Proxy (ZqBasic q Int64))) $ \pl ->
        c_liftRq pl (castPtr ps) (castPtr pd) 1 nullPtr >>= check "lolb_liftRq"
    return (GT' dst)
  _ -> error "liftT: unreachable"

reduceT :: forall m q . (Fact m, Reflects q Int64) => GT m Int64 -> GT m (ZqBasic q Int64)
reduceT v = case toGT v of
  GT (GT' src) -> GT $ unsafePerformIO $ do
    dst <- newDev (totM (Proxy :: Proxy m))
    withForeignPtr src $ \ps -> withForeignPtr dst $ \pd ->
      withPlan (proxy (planFor (Proxy :: Proxy m)) (Proxy :: Proxy (ZqBasic q Int64))) $ \pl ->
        c_reduceRq pl (castPtr ps) 1 (castPtr pd) 1 nullPtr >>= check "lolb_reduceRq"
    return (GT' dst)
  _ -> error "reduceT: unreachable"

-- | @rescalePow :: Rescale (a, b) b@ (Prelude.hs:226-265): drop the first limb
rescaleDropT :: forall m a b . (Fact m, DevElt (a, b), DevElt b) => GT m (a, b) -> GT m b
rescaleDropT v = case toGT v of
  GT (GT' src) -> GT $ unsafePerformIO $ do
    dst <- newDev (totM (Proxy :: Proxy m))
    withForeignPtr src $ \ps -> withForeignPtr dst $ \pd ->
      withPlan (proxy (planFor (Proxy :: Proxy m)) (Proxy :: Proxy (a, b))) $ \pl ->
        c_rescaleDropRq pl 0 (castPtr ps) (castPtr pd) 1 nullPtr >>= check "lolb_rescaleDropRq"
    return (GT' dst)
  _ -> error "rescaleDropT: unreachable"

-- ------------------------------------------------------------------------------------------------ host batches

-- | Many host-resident ring elements through ONE call of the library's pipelined host path (lolb_rq_apply_host: chunked
-- H2D -> kernels -> D2H): what replaces one `SV.thaw` + one FFI call per element (CPP.hs:325-337) when the data
-- must stay on the host.
batchRq :: forall m q . (Fact m, Reflects q Int64)
        => String -> Proxy m -> [SV.Vector (ZqBasic q Int64)] -> [SV.Vector (ZqBasic q Int64)]
batchRq ops pm xs = unsafePerformIO $ do
  let n = totM pm
  buf <- SV.thaw (SV.concat xs)
  withPlan (proxy (planFor pm) (Proxy :: Proxy (ZqBasic q Int64))) $ \pl ->
    SM.unsafeWith buf $ \p -> applyHostRq pl ops (castPtr p) (fromIntegral (length xs))
  out <- SV.unsafeFreeze buf
  return [ SV.slice (i * n) n out | i <- [0 .. length xs - 1] ]

crtBatch, crtInvBatch :: (Fact m, Reflects q Int64) => Proxy m -> [SV.Vector (ZqBasic q Int64)] -> [SV.Vector (ZqBasic q Int64)]
crtBatch    = batchRq "CRT"
crtInvBatch = batchRq "CRTInv"
