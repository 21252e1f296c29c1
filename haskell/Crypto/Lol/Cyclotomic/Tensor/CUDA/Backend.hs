{-|
Module      : Crypto.Lol.Cyclotomic.Tensor.CUDA.Backend
Description : FFI to libctensor_b200.so (include/lol_b200.h): plans and batched, device-backed operators.

NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no GHC/stack/cabal here (SURVEY.md, fact 2), so this module
is written against lol-0.7.0.0 / lol-cpp-0.0.0.4 and has never been type-checked.  It mirrors
lol-cpp/Crypto/Lol/Cyclotomic/Tensor/CPP/Backend.hs: the per-element `Dispatch'` class of that file keeps
working unchanged against the drop-in symbols of the new library (same names, same C signatures), and this module
adds what the C++ back end never had -- a plan handle and whole-batch calls.
-}

{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE ScopedTypeVariables      #-}

module Crypto.Lol.Cyclotomic.Tensor.CUDA.Backend
( Plan, withPlanRq, applyHostRq
, lolbDeviceAvailable, lolbLastError
  -- * device-resident SymmSHE steps (raw imports; operands are device addresses)
, DevPtr, c_crtRq, c_crtInvRq, c_ctMulRq, c_gadgetLength, c_decomposeRq, c_decomposeCrtRq, c_knapsackRq
  -- * ring extensions O_m'/O_m (raw imports; one ExtStruct per '(m, m'), Tensor.hs:380-498 / CPP/Extension.hs:54-129)
, ExtStruct, ringRq, ringInt, ringDouble, ringComplex
, c_extCreate, p_extDestroy, c_twacePowDec, c_embedPow, c_embedDec, c_embedCRT, c_coeffsPowDec, c_twaceCRT
  -- * coefficient-wise maps (fmapT lift / reduce / rescale, UCyc.hs:267-300; roundCoset, Prelude.hs:155-162)
, c_liftRq, c_reduceRq, c_rescaleDropRq, c_rescaleModRq, c_roundCosetRq
) where

import Control.Exception      (bracket, throwIO, ErrorCall (..))
import Control.Monad          (when)
import Data.Int
import Foreign.C.String       (CString, peekCString, withCString)
import Foreign.ForeignPtr
import Foreign.Marshal.Alloc  (alloca)
import Foreign.Marshal.Array  (withArray, withArrayLen)
import Foreign.Ptr
import Foreign.Storable       (peek)

-- | Same C representation of a prime power as CPP/Backend.hs:78.
type CPP = (Int16, Int16)

-- | Opaque @lolb_plan@.
data PlanStruct
newtype Plan = Plan (ForeignPtr PlanStruct)

foreign import ccall unsafe "lolb_plan_create_rq" c_planCreateRq ::
  Ptr (Ptr PlanStruct) -> Ptr CPP -> Int16 -> Int16 -> Ptr Int64
  -> Ptr (Ptr Int64) -> Ptr (Ptr Int64) -> Ptr Int64 -> IO Int32
foreign import ccall unsafe "&lolb_plan_destroy" p_planDestroy :: FunPtr (Ptr PlanStruct -> IO ())
foreign import ccall unsafe "lolb_rq_apply_host" c_applyHostRq ::
  Ptr PlanStruct -> CString -> Ptr Int64 -> Int64 -> IO Int32
-- | Device-resident batches ([batch][n][tupSize] Int64 in GPU memory; allocation is the caller's, e.g. cudaMalloc).
-- The SymmSHE steps between the CRTs (lol-apps SymmSHE.hs:302-314, 359-372, 443-449) on such batches:
type DevPtr a = Ptr a
foreign import ccall unsafe "lolb_tensorCRTRq" c_crtRq :: Ptr PlanStruct -> DevPtr Int64 -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_tensorCRTInvRq" c_crtInvRq :: Ptr PlanStruct -> DevPtr Int64 -> Int64 -> Ptr () -> IO Int32
-- (d0,d1,d2) <- mulG <$> [a0,a1] * [b0,b1]
foreign import ccall unsafe "lolb_ctMulRq" c_ctMulRq ::
  Ptr PlanStruct -> DevPtr Int64 -> DevPtr Int64 -> DevPtr Int64 -> DevPtr Int64
  -> DevPtr Int64 -> DevPtr Int64 -> DevPtr Int64 -> Int64 -> Int32 -> Ptr () -> IO Int32
-- number of gadget digits; base 0 = TrivGad, b >= 2 = BaseBGad b
foreign import ccall unsafe "lolb_gadgetLength" c_gadgetLength :: Ptr PlanStruct -> Int64 -> IO Int32
-- digits <- fmap reduce <$> decompose x          (x in the powerful basis)
foreign import ccall unsafe "lolb_decomposeRq" c_decomposeRq ::
  Ptr PlanStruct -> DevPtr Int64 -> DevPtr Int64 -> Int64 -> Int64 -> Ptr () -> IO Int32
-- digits <- adviseCRT <$> (fmap reduce <$> decompose x)
foreign import ccall unsafe "lolb_decomposeCrtRq" c_decomposeCrtRq ::
  Ptr PlanStruct -> DevPtr Int64 -> DevPtr Int64 -> Int64 -> Int64 -> Ptr () -> IO Int32
-- [c0,c1] += knapsack hint digits
foreign import ccall unsafe "lolb_knapsackRq" c_knapsackRq ::
  Ptr PlanStruct -> DevPtr Int64 -> Int32 -> DevPtr Int64 -> DevPtr Int64 -> DevPtr Int64 -> Int64 -> Ptr () -> IO Int32

-- | Opaque @lolb_ext@: the index tables of one extension O_m'/O_m on the device.
data ExtStruct
-- | The @ring@ argument of the extension operators (LOLB_RING_* in lol_b200.h).
ringRq, ringInt, ringDouble, ringComplex :: Int32
ringRq = 0; ringInt = 1; ringDouble = 2; ringComplex = 3

foreign import ccall unsafe "lolb_ext_create" c_extCreate ::
  Ptr (Ptr ExtStruct) -> Ptr PlanStruct -> Ptr PlanStruct -> IO Int32
foreign import ccall unsafe "&lolb_ext_destroy" p_extDestroy :: FunPtr (Ptr ExtStruct -> IO ())
-- twacePowDec' (Extension.hs:99-103): O_m' -> O_m
foreign import ccall unsafe "lolb_twacePowDec" c_twacePowDec ::
  Ptr ExtStruct -> Int32 -> DevPtr a -> DevPtr a -> Int64 -> Ptr () -> IO Int32
-- embedPow', embedDec', embedCRT' (Extension.hs:60-85): O_m -> O_m'
foreign import ccall unsafe "lolb_embedPow" c_embedPow ::
  Ptr ExtStruct -> Int32 -> DevPtr a -> DevPtr a -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_embedDec" c_embedDec ::
  Ptr ExtStruct -> Int32 -> DevPtr a -> DevPtr a -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_embedCRT" c_embedCRT ::
  Ptr ExtStruct -> Int32 -> DevPtr a -> DevPtr a -> Int64 -> Ptr () -> IO Int32
-- coeffs' (Extension.hs:90-93): the phi'/phi output elements are laid out back to back per input element
foreign import ccall unsafe "lolb_coeffsPowDec" c_coeffsPowDec ::
  Ptr ExtStruct -> Int32 -> DevPtr a -> DevPtr a -> Int64 -> Ptr () -> IO Int32
-- twaceCRT' (Extension.hs:110-129); status 2 (LOLB_ERR_NO_CRT) is the reference's Nothing
foreign import ccall unsafe "lolb_twaceCRT" c_twaceCRT ::
  Ptr ExtStruct -> Int32 -> DevPtr a -> DevPtr a -> Int64 -> Ptr () -> IO Int32

-- fmapT lift / fmapT reduce (UCyc.hs:267-296); reduce takes 1 or tupSize integers per coefficient
foreign import ccall unsafe "lolb_liftRq" c_liftRq ::
  Ptr PlanStruct -> DevPtr Int64 -> DevPtr Int64 -> Int64 -> Ptr () -> IO Int32
foreign import ccall unsafe "lolb_reduceRq" c_reduceRq ::
  Ptr PlanStruct -> DevPtr Int64 -> Int32 -> DevPtr Int64 -> Int64 -> Ptr () -> IO Int32
-- rescalePow over Rescale (a,b) b (drop = 0) / Rescale (a,b) a (drop = tupSize-1), Prelude.hs:226-265
foreign import ccall unsafe "lolb_rescaleDropRq" c_rescaleDropRq ::
  Ptr PlanStruct -> Int32 -> DevPtr Int64 -> DevPtr Int64 -> Int64 -> Ptr () -> IO Int32
-- fmapT rescaleMod (Prelude.hs:143-153); second argument: the target moduli (host pointer)
foreign import ccall unsafe "lolb_rescaleModRq" c_rescaleModRq ::
  Ptr PlanStruct -> Ptr Int64 -> DevPtr Int64 -> DevPtr Int64 -> Int64 -> Ptr () -> IO Int32
-- roundCoset <$> zp <*> e (Prelude.hs:155-162); a null coset pointer rounds to the nearest integer (errorRounded)
foreign import ccall unsafe "lolb_roundCosetRq" c_roundCosetRq ::
  Ptr PlanStruct -> DevPtr Double -> DevPtr Int64 -> DevPtr Int64 -> Int64 -> Ptr () -> IO Int32

foreign import ccall unsafe "lolb_last_error" c_lastError :: IO CString
foreign import ccall unsafe "lolb_device_available" c_deviceAvailable :: IO Int32

lolbLastError :: IO String
lolbLastError = c_lastError >>= peekCString

lolbDeviceAvailable :: IO Bool
lolbDeviceAvailable = (/= 0) <$> c_deviceAvailable

check :: String -> Int32 -> IO ()
check what st = when (st /= 0) $ do
  msg <- lolbLastError
  throwIO $ ErrorCall $ what ++ ": libctensor_b200 status " ++ show st ++ ": " ++ msg

-- | Plan for index @m = prod pps@ over the moduli @qs@; root tables are derived inside the library exactly as
-- 'Crypto.Lol.Types.Unsafe.ZqBasic' derives them (smallest generator of Z_q^*), so results agree with 'CT'.
withPlanRq :: [CPP] -> [Int64] -> (Plan -> IO a) -> IO a
withPlanRq pps qs act =
  withArrayLen pps $ \npe ppe ->
  withArrayLen qs $ \k pqs ->
  alloca $ \(pp :: Ptr (Ptr PlanStruct)) -> do
    c_planCreateRq pp ppe (fromIntegral npe) (fromIntegral k) pqs nullPtr nullPtr nullPtr >>= check "lolb_plan_create_rq"
    raw <- peek pp
    fp <- newForeignPtr p_planDestroy raw
    act (Plan fp)

-- | @applyHostRq plan "CRT,CRTInv" buf batch@: transform @batch@ ring elements laid out back to back in @buf@
-- (the element layout of CPP/Backend.hs:80-90) in place; the library pipelines host->device copy, kernels and
-- device->host copy over chunks.  Operator names: CRT, CRTInv, L, LInv, GPow, GDec, GInvPow, GInvDec, MulGCRT, DivGCRT.
applyHostRq :: Plan -> String -> Ptr Int64 -> Int64 -> IO ()
applyHostRq (Plan fp) ops buf batch =
  withForeignPtr fp $ \p -> withCString ops $ \cops ->
    c_applyHostRq p cops buf batch >>= check "lolb_rq_apply_host"
