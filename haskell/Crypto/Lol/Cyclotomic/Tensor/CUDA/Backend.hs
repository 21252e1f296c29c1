{-|
Module      : Crypto.Lol.Cyclotomic.Tensor.CUDA.Backend
Description : FFI to libctensor_b200.so (include/lol_b200.h): plans, device memory and the batched, device-resident operators.

NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no GHC/stack/cabal here (SURVEY.md, fact 2), so this module is written
against lol-0.7.0.0 / lol-cpp-0.0.0.4 and has never been type-checked.  It is the counterpart of
lol-cpp/Crypto/Lol/Cyclotomic/Tensor/CPP/Backend.hs:304-337 for the device-resident 'GT' type of
Crypto.Lol.Cyclotomic.Tensor.CUDA: where that file imports 29 per-element host-pointer symbols, this one imports the
plan-based `lolb_*` entry points, whose operands are DEVICE pointers (every C signature is in include/lol_b200.h).

Note on "pinned": Haskell 'Storable' vectors are pinned against the GARBAGE COLLECTOR (their address is stable during a
foreign call), not page-locked for DMA; `lolb_rq_apply_host` therefore stages such buffers through the driver.  Callers
that want true asynchronous copies allocate with `lolb_host_alloc` (cudaHostAlloc).
-}

{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE ScopedTypeVariables      #-}

module Crypto.Lol.Cyclotomic.Tensor.CUDA.Backend
( CPP, Plan, PlanStruct, ExtStruct
, newPlanRq, newPlanC, withPlan, withExt, check
, lolbDeviceAvailable, lolbLastError
  -- * device memory
, devAlloc, p_devFree, devUpload, devDownload, devCopy
  -- * operator shapes
, DevOp, DevOp2, DevOpOut, DevOpStatus, ExtOp, noOp, noOpS
  -- * single-index operators (device pointers)
, c_lRq, c_lInvRq, c_gPowRq, c_gDecRq, c_gInvPowRq, c_gInvDecRq, c_crtRq, c_crtInvRq, c_mulRq
, c_lR, c_lInvR, c_gPowR, c_gDecR, c_gInvPowR, c_gInvDecR, c_normSqR
, c_lDouble, c_lInvDouble, c_normSqD
, c_lC, c_lInvC, c_gPowC, c_gDecC, c_gInvPowC, c_gInvDecC, c_crtC, c_crtInvC, c_mulC
, c_planGcrtDev, c_tGaussianDec
  -- * SymmSHE steps between the CRTs (SymmSHE.hs:302-314, 359-372, 443-449)
, c_ctMulRq, c_gadgetLength, c_decomposeRq, c_decomposeCrtRq, c_knapsackRq
  -- * ring extensions (CPP/Extension.hs:54-143)
, ringRq, ringInt, ringDouble, ringComplex
, c_twacePowDec, c_embedPow, c_embedDec, c_embedCRT, c_coeffsPowDec, c_twaceCRT, c_powBasisPow
  -- * coefficient-wise maps
, c_liftRq, c_reduceRq, c_rescaleDropRq, c_rescaleModRq, c_roundCosetRq
  -- * host batches
, applyHostRq
) where

import Control.Exception      (throwIO, ErrorCall (..))
import Control.Monad          (when)
import Data.Int
import Data.IORef
import qualified Data.Map.Strict as M
import Data.Word
import Foreign.C.String       (CString, peekCString, withCString)
import Foreign.ForeignPtr
import Foreign.Marshal.Alloc  (alloca)
import Foreign.Marshal.Array  (withArrayLen)
import Foreign.Ptr
import Foreign.Storable
import System.IO.Unsafe       (unsafePerformIO)

-- | C @PrimeExponent@ (types.h:27-31): two int16.  CPP/Backend.hs:78 uses a tuple with an orphan 'Storable' instance;
-- a newtype with its own instance keeps this module self-contained.
newtype PE = PE (Int16, Int16)
instance Storable PE where
  sizeOf _    = 4
  alignment _ = 2
  peek p      = do a <- peekByteOff p 0; b <- peekByteOff p 2; return (PE (a, b))
  poke p (PE (a, b)) = pokeByteOff p 0 a >> pokeByteOff p 2 b
type CPP = (Int16, Int16)

data PlanStruct
data ExtStruct
-- | a plan owns device tables; destroyed by the GC through lolb_plan_destroy
newtype Plan = Plan (ForeignPtr PlanStruct)

foreign import ccall unsafe "lolb_plan_create_rq" c_planCreateRq ::
  Ptr (Ptr PlanStruct) -> Ptr PE -> Int16 -> Int16 -> Ptr Int64 -> Ptr (Ptr Int64) -> Ptr (Ptr Int64) -> Ptr Int64 -> IO Int32
foreign import ccall unsafe "lolb_plan_create_c" c_planCreateC ::
  Ptr (Ptr PlanStruct) -> Ptr PE -> Int16 -> Int16 -> IO Int32
foreign import ccall unsafe "&lolb_plan_destroy" p_planDestroy :: FunPtr (Ptr PlanStruct -> IO ())
foreign import ccall unsafe "lolb_plan_gcrt_dev" c_planGcrtDev :: Ptr PlanStruct -> Int32 -> IO (Ptr Int64)
foreign import ccall unsafe "lolb_last_error" c_lastError :: IO CString
foreign import ccall unsafe "lolb_device_available" c_deviceAvailable :: IO Int32

lolbLastError :: IO String
lolbLastError = c_lastError >>= peekCString

lolbDeviceAvailable :: IO Bool
lolbDeviceAvailable = (/= 0) <$> c_deviceAvailable

-- | status 0 = LOLB_OK; anything else carries a message (lolb_last_error)
check :: String -> Int32 -> IO ()
check what st = when (st /= 0) $ do
  msg <- lolbLastError
  throwIO $ ErrorCall $ what ++ ": libctensor_b200 status " ++ show st ++ ": " ++ msg

-- | Plan for index @m = prod pps@ over the moduli @qs@; root tables are derived inside the library exactly as
-- 'Crypto.Lol.Types.Unsafe.ZqBasic' derives them (smallest generator of Z_q^*, ZqBasic.hs:144-165), so results agree with 'CT'.
newPlanRq :: [CPP] -> [Int64] -> IO Plan
newPlanRq pps qs =
  withArrayLen (map PE pps) $ \npe ppe ->
  withArrayLen qs $ \k pqs ->
  alloca $ \pp -> do
    c_planCreateRq pp ppe (fromIntegral npe) (fromIntegral k) pqs nullPtr nullPtr nullPtr >>= check "lolb_plan_create_rq"
    Plan <$> (peek pp >>= newForeignPtr p_planDestroy)

-- | Modulus-free plan (Int64 / Double / Complex Double elements), @k@ = tuple size
newPlanC :: [CPP] -> Int -> IO Plan
newPlanC pps k =
  withArrayLen (map PE pps) $ \npe ppe ->
  alloca $ \pp -> do
    c_planCreateC pp ppe (fromIntegral npe) (fromIntegral k) >>= check "lolb_plan_create_c"
    Plan <$> (peek pp >>= newForeignPtr p_planDestroy)

withPlan :: Plan -> (Ptr PlanStruct -> IO a) -> IO a
withPlan (Plan fp) = withForeignPtr fp

-- ------------------------------------------------------------------------------------------------ device memory
foreign import ccall unsafe "lolb_dev_alloc"    c_devAlloc    :: Word64 -> IO (Ptr ())
foreign import ccall unsafe "&lolb_dev_free"    p_devFree     :: FunPtr (Ptr a -> IO ())
foreign import ccall unsafe "lolb_dev_upload"   c_devUpload   :: Ptr () -> Ptr () -> Word64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_dev_download" c_devDownload :: Ptr () -> Ptr () -> Word64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_dev_copy"     c_devCopy     :: Ptr () -> Ptr () -> Word64 -> Ptr () -> IO Int32

devAlloc :: Word64 -> IO (Ptr ())
devAlloc bytes = do
  p <- c_devAlloc bytes
  when (p == nullPtr) $ check "lolb_dev_alloc" 1
  return p
devUpload, devDownload, devCopy :: Ptr () -> Ptr () -> Word64 -> IO ()
devUpload   d s n = c_devUpload   d s n nullPtr >>= check "lolb_dev_upload"
devDownload d s n = c_devDownload d s n nullPtr >>= check "lolb_dev_download"
devCopy     d s n = c_devCopy     d s n nullPtr >>= check "lolb_dev_copy"

-- ------------------------------------------------------------------------------------------------ operator shapes
-- | in-place operator on @batch@ device-resident elements: plan, data, batch, CUDA stream (NULL = default stream)
type DevOp       = Ptr PlanStruct -> Ptr () -> Int64 -> Ptr () -> IO Int32
-- | y <- y . b (b_batch = 1 broadcasts one element)
type DevOp2      = Ptr PlanStruct -> Ptr () -> Ptr () -> Int64 -> Int64 -> Ptr () -> IO Int32
-- | reduction to one value per element
type DevOpOut    = Ptr PlanStruct -> Ptr () -> Ptr () -> Int64 -> Ptr () -> IO Int32
-- | divG: True = divisible (CPP.hs:321-323)
type DevOpStatus = Ptr PlanStruct -> Ptr () -> IO Bool
type ExtOp       = Ptr ExtStruct -> Int32 -> Ptr () -> Ptr () -> Int64 -> Ptr () -> IO Int32

noOp :: DevOp
noOp _ _ _ _ = return 1         -- Backend.hs:267-283: the (type, operator) pair does not exist in the reference either
noOpS :: DevOpStatus
noOpS _ _ = return False

foreign import ccall unsafe "lolb_tensorLRq"      c_lRq      :: DevOp
foreign import ccall unsafe "lolb_tensorLInvRq"   c_lInvRq   :: DevOp
foreign import ccall unsafe "lolb_tensorGPowRq"   c_gPowRq   :: DevOp
foreign import ccall unsafe "lolb_tensorGDecRq"   c_gDecRq   :: DevOp
foreign import ccall unsafe "lolb_tensorGInvPowRq" c_gInvPowRq' :: DevOp
foreign import ccall unsafe "lolb_tensorGInvDecRq" c_gInvDecRq' :: DevOp
foreign import ccall unsafe "lolb_tensorCRTRq"    c_crtRq    :: DevOp
foreign import ccall unsafe "lolb_tensorCRTInvRq" c_crtInvRq :: DevOp
foreign import ccall unsafe "lolb_mulRq"          c_mulRq    :: DevOp2
foreign import ccall unsafe "lolb_tensorLR"       c_lR       :: DevOp
foreign import ccall unsafe "lolb_tensorLInvR"    c_lInvR    :: DevOp
foreign import ccall unsafe "lolb_tensorGPowR"    c_gPowR    :: DevOp
foreign import ccall unsafe "lolb_tensorGDecR"    c_gDecR    :: DevOp
foreign import ccall unsafe "lolb_tensorGInvPowR" c_gInvPowR' :: Ptr PlanStruct -> Ptr () -> Ptr Int16 -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_tensorGInvDecR" c_gInvDecR' :: Ptr PlanStruct -> Ptr () -> Ptr Int16 -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_tensorNormSqR"  c_normSqR  :: DevOpOut
foreign import ccall unsafe "lolb_tensorLDouble"    c_lDouble    :: DevOp
foreign import ccall unsafe "lolb_tensorLInvDouble" c_lInvDouble :: DevOp
foreign import ccall unsafe "lolb_tensorNormSqD"  c_normSqD  :: DevOpOut
foreign import ccall unsafe "lolb_tensorLC"       c_lC       :: DevOp
foreign import ccall unsafe "lolb_tensorLInvC"    c_lInvC    :: DevOp
foreign import ccall unsafe "lolb_tensorGPowC"    c_gPowC    :: DevOp
foreign import ccall unsafe "lolb_tensorGDecC"    c_gDecC    :: DevOp
foreign import ccall unsafe "lolb_tensorGInvPowC" c_gInvPowC' :: DevOp
foreign import ccall unsafe "lolb_tensorGInvDecC" c_gInvDecC' :: DevOp
foreign import ccall unsafe "lolb_tensorCRTC"     c_crtC     :: DevOp
foreign import ccall unsafe "lolb_tensorCRTInvC"  c_crtInvC  :: DevOp
foreign import ccall unsafe "lolb_mulC"           c_mulC     :: DevOp2
-- tGaussianDec v on the device: plan, v, seed, first element of the stream, output, batch, CUDA stream
foreign import ccall unsafe "lolb_tGaussianDec" c_tGaussianDec ::
  Ptr PlanStruct -> Double -> Word64 -> Word64 -> Ptr () -> Int64 -> Ptr () -> IO Int32

-- Rq / C: status LOLB_ERR_NOT_INVERTIBLE (4) means rad_odd(m) is not a unit -> Nothing (g.cpp:196-198)
statusOf :: DevOp -> DevOpStatus
statusOf op pl y = do st <- op pl y 1 nullPtr
                      if st == 0 then return True else if st == 4 then return False else check "lolb_tensorGInv*" st >> return False
c_gInvPowRq, c_gInvDecRq, c_gInvPowC, c_gInvDecC :: DevOpStatus
c_gInvPowRq = statusOf c_gInvPowRq'; c_gInvDecRq = statusOf c_gInvDecRq'
c_gInvPowC  = statusOf c_gInvPowC';  c_gInvDecC  = statusOf c_gInvDecC'

-- Z: the per-element divisibility verdict comes back through a device int16 (g.cpp:169-184, 222-237)
statusOfR :: (Ptr PlanStruct -> Ptr () -> Ptr Int16 -> Int64 -> Ptr () -> IO Int32) -> DevOpStatus
statusOfR op pl y = do
  okDev <- devAlloc 8
  op pl y (castPtr okDev) 1 nullPtr >>= check "lolb_tensorGInv*R"
  ok <- alloca $ \(h :: Ptr Int16) -> devDownload (castPtr h) okDev 2 >> peek h
  c_devFreeNow okDev
  return (ok /= 0)
foreign import ccall unsafe "lolb_dev_free" c_devFreeNow :: Ptr () -> IO ()
c_gInvPowR, c_gInvDecR :: DevOpStatus
c_gInvPowR = statusOfR c_gInvPowR'; c_gInvDecR = statusOfR c_gInvDecR'

-- ------------------------------------------------------------------------------------------------ SymmSHE steps
foreign import ccall unsafe "lolb_ctMulRq" c_ctMulRq ::
  Ptr PlanStruct -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> Int64 -> Int32 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_gadgetLength" c_gadgetLength :: Ptr PlanStruct -> Int64 -> IO Int32
foreign import ccall unsafe "lolb_decomposeRq" c_decomposeRq ::
  Ptr PlanStruct -> Ptr Int64 -> Ptr Int64 -> Int64 -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_decomposeCrtRq" c_decomposeCrtRq ::
  Ptr PlanStruct -> Ptr Int64 -> Ptr Int64 -> Int64 -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_knapsackRq" c_knapsackRq ::
  Ptr PlanStruct -> Ptr Int64 -> Int32 -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> Int64 -> Ptr () -> IO Int32

-- ------------------------------------------------------------------------------------------------ ring extensions
ringRq, ringInt, ringDouble, ringComplex :: Int32
ringRq = 0; ringInt = 1; ringDouble = 2; ringComplex = 3

foreign import ccall unsafe "lolb_ext_create" c_extCreate :: Ptr (Ptr ExtStruct) -> Ptr PlanStruct -> Ptr PlanStruct -> IO Int32
foreign import ccall unsafe "&lolb_ext_destroy" p_extDestroy :: FunPtr (Ptr ExtStruct -> IO ())
foreign import ccall unsafe "lolb_twacePowDec"  c_twacePowDec  :: ExtOp
foreign import ccall unsafe "lolb_embedPow"     c_embedPow     :: ExtOp
foreign import ccall unsafe "lolb_embedDec"     c_embedDec     :: ExtOp
foreign import ccall unsafe "lolb_embedCRT"     c_embedCRT     :: ExtOp
foreign import ccall unsafe "lolb_coeffsPowDec" c_coeffsPowDec :: ExtOp
foreign import ccall unsafe "lolb_twaceCRT"     c_twaceCRT     :: ExtOp
foreign import ccall unsafe "lolb_powBasisPow"  c_powBasisPow  :: Ptr ExtStruct -> Int32 -> Ptr () -> Ptr () -> IO Int32

{-# NOINLINE extCache #-}
extCache :: IORef (M.Map (Ptr PlanStruct, Ptr PlanStruct) (ForeignPtr ExtStruct, Plan, Plan))
extCache = unsafePerformIO $ newIORef M.empty

-- | the extension O_m'/O_m of two plans (index tables on the device), created once and kept with its plans
withExt :: Plan -> Plan -> (Ptr ExtStruct -> IO a) -> IO a
withExt lo@(Plan flo) hi@(Plan fhi) act =
  withForeignPtr flo $ \plo -> withForeignPtr fhi $ \phi -> do
    cache <- readIORef extCache
    fe <- case M.lookup (plo, phi) cache of
      Just (e, _, _) -> return e
      Nothing -> alloca $ \pp -> do
        c_extCreate pp plo phi >>= check "lolb_ext_create"
        e <- peek pp >>= newForeignPtr p_extDestroy
        atomicModifyIORef' extCache (\c -> (M.insert (plo, phi) (e, lo, hi) c, ()))
        return e
    withForeignPtr fe act

-- ------------------------------------------------------------------------------------------------ coefficient-wise maps
foreign import ccall unsafe "lolb_liftRq" c_liftRq :: Ptr PlanStruct -> Ptr Int64 -> Ptr Int64 -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_reduceRq" c_reduceRq :: Ptr PlanStruct -> Ptr Int64 -> Int32 -> Ptr Int64 -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_rescaleDropRq" c_rescaleDropRq :: Ptr PlanStruct -> Int32 -> Ptr Int64 -> Ptr Int64 -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_rescaleModRq" c_rescaleModRq :: Ptr PlanStruct -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_roundCosetRq" c_roundCosetRq :: Ptr PlanStruct -> Ptr Double -> Ptr Int64 -> Ptr Int64 -> Int64 -> Ptr () -> IO Int32

-- ------------------------------------------------------------------------------------------------ host batches
foreign import ccall unsafe "lolb_rq_apply_host" c_applyHostRq :: Ptr PlanStruct -> CString -> Ptr Int64 -> Int64 -> IO Int32

-- | @applyHostRq plan "CRT,CRTInv" buf batch@: transform @batch@ ring elements laid out back to back in host memory, in place
-- (chunked H2D -> kernels -> D2H).  Operator names: CRT, CRTInv, L, LInv, GPow, GDec, GInvPow, GInvDec, MulGCRT, DivGCRT.
applyHostRq :: Ptr PlanStruct -> String -> Ptr Int64 -> Int64 -> IO ()
applyHostRq p ops buf batch = withCString ops $ \cops -> c_applyHostRq p cops buf batch >>= check "lolb_rq_apply_host"
