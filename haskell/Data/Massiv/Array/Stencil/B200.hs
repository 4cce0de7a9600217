{-# LANGUAGE BangPatterns #-}
{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE LambdaCase #-}
{-# LANGUAGE RecordWildCards #-}
{-# LANGUAGE ScopedTypeVariables #-}
-- |
-- Module      : Data.Massiv.Array.Stencil.B200
-- Description : B200 device target for massiv's stencil / windowed-load path
--
-- Drop-in companion of "Data.Massiv.Array.Stencil".  massiv's own 'Stencil' is an opaque closure
-- (Data.Massiv.Array.Stencil.Internal, lines 30-34) and 'Comp' is a closed type owned by the
-- @scheduler@ package, so a device target cannot be a new 'Comp' constructor: it is this module.
-- Every constructor here mirrors the one of the same name in massiv but returns a /reified/
-- stencil ('StencilB200') that lowers to @mb_stencil_desc@ (include/massiv_b200.h); 'computeB200'
-- replaces @compute@ of the @Array DW@ that 'mapStencil' would have produced.  Closure stencils
-- ('makeStencil') have no device form: they simply are not constructible here, there is no
-- silent CPU fallback.
--
-- NOTE: this file is shipped as source only.  The build image has no GHC (SURVEY.md, fact 2), so it
-- has not been compiled; INTEGRATION.md lists the steps a maintainer runs.  The C ABI it binds is
-- exercised by tests/ (through ctypes and through a plain C harness, tests/cabi_harness.c).
module Data.Massiv.Array.Stencil.B200
  ( -- * Device
    DeviceB200
  , withDeviceB200
    -- * Reified stencils (same names as Data.Massiv.Array.Stencil)
  , StencilB200
  , idStencil
  , sumStencil
  , productStencil
  , avgStencil
  , maxStencil
  , minStencil
  , makeConvolutionStencilFromKernel
  , makeCorrelationStencilFromKernel
  , lifeStencil
  , sobelMagnitude
  , assumeFinite
    -- * Integral approximation (Data.Massiv.Array.Numeric.Integral)
  , midpointStencil
  , trapezoidStencil
  , simpsonsStencil
  , integrateWithB200
    -- * Application
  , computeB200
  , mapStencilB200
  , applyStencilB200
  , computeWithStrideB200
  , iterateStencilB200
    -- * Errors
  , B200Exception(..)
  ) where

import Control.Exception (Exception, bracket, throwIO)
import Control.Monad (when)
import Data.Massiv.Array
  ( Array, Border(..), Ix2, Index, Padding(..), S, Stride, Sz(..), size, unStride )
import qualified Data.Massiv.Array as A
import qualified Data.Massiv.Array.Unsafe as A
import Data.Massiv.Core.Index (IndexException(..), SizeException(..))
import Data.Int
import Data.Word
import Foreign
import Foreign.C.String (CString, peekCString)
import Foreign.C.Types

-- ---------------------------------------------------------------------------------------------
-- C ABI (include/massiv_b200.h).  All calls are `safe`: they last milliseconds to seconds and an
-- `unsafe` call would stall every capability and block GC.
-- ---------------------------------------------------------------------------------------------

data MbCtx

foreign import ccall safe "mb_ctx_create"  c_mb_ctx_create  :: CInt -> Ptr (Ptr MbCtx) -> IO CInt
foreign import ccall safe "mb_ctx_destroy" c_mb_ctx_destroy :: Ptr MbCtx -> IO CInt
foreign import ccall safe "mb_last_error"  c_mb_last_error  :: Ptr MbCtx -> IO CString
foreign import ccall safe "mb_out_dims"    c_mb_out_dims    :: Ptr MbDesc -> Ptr Int64 -> Ptr Int64 -> IO CInt
foreign import ccall safe "mb_run"         c_mb_run
  :: Ptr MbCtx -> Ptr MbDesc -> Ptr Int64 -> Ptr () -> Ptr () -> IO CInt
foreign import ccall safe "mb_iterate"     c_mb_iterate
  :: Ptr MbCtx -> Ptr MbDesc -> Ptr Int64 -> Ptr () -> Ptr () -> Int64 -> IO CInt

-- | @mb_stencil_desc@: 4 x int32, 5 x 5 x int64, 8 fill bytes, 2 pointers, int64, pointer, 2 x uint32
data MbDesc

mbDescBytes :: Int
mbDescBytes = 16 + 5 * 5 * 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8  -- ... flags, n_post (4 + 4), post pointer: 272

-- ---------------------------------------------------------------------------------------------

data B200Exception = B200Error !Int String
  deriving Show

instance Exception B200Exception

newtype DeviceB200 = DeviceB200 (Ptr MbCtx)

-- | One context = one GPU (the analogue of the scheduler 'withMassivScheduler_' picks,
-- Data.Massiv.Core.Loop lines 468-474).
withDeviceB200 :: Int -> (DeviceB200 -> IO a) -> IO a
withDeviceB200 ordinal = bracket open close
  where
    open = alloca $ \pp -> do
      st <- c_mb_ctx_create (fromIntegral ordinal) pp
      when (st /= 0) $ throwIO (B200Error (fromIntegral st) "mb_ctx_create: no usable CUDA device (no CPU fallback)")
      DeviceB200 <$> peek pp
    close (DeviceB200 p) = () <$ c_mb_ctx_destroy p

data Kind = KId | KSum | KProduct | KAvg | KMax | KMin | KConv | KCorr | KTaps | KLife | KMag2
          | KIntMidpoint | KIntTrapezoid | KIntSimpson
  deriving (Eq, Enum)

-- | Element types with a device representation (Storable sizes, Bool is 4 bytes).
class Storable e => ElemB200 e where
  elemCode :: proxy e -> Int32

instance ElemB200 Word8  where elemCode _ = 0
instance ElemB200 Int8   where elemCode _ = 1
instance ElemB200 Word16 where elemCode _ = 2
instance ElemB200 Int16  where elemCode _ = 3
instance ElemB200 Word32 where elemCode _ = 4
instance ElemB200 Int32  where elemCode _ = 5
instance ElemB200 Word64 where elemCode _ = 6
instance ElemB200 Int64  where elemCode _ = 7
instance ElemB200 Int    where elemCode _ = 7
instance ElemB200 Float  where elemCode _ = 8
instance ElemB200 Double where elemCode _ = 9
instance ElemB200 Bool   where elemCode _ = 10

-- | The reified counterpart of @Stencil ix e e@.
data StencilB200 ix e = StencilB200
  { sbKind   :: !Kind
  , sbSize   :: !(Sz ix)
  , sbCenter :: !ix
  , sbCoeffs :: Maybe (Array S ix e)
  , sbCoeffs2 :: Maybe (Array S ix e)
  , sbFlags  :: !Word32
  }

foldStencilB200 :: Index ix => Kind -> Sz ix -> StencilB200 ix e
foldStencilB200 k sz = StencilB200 k sz A.zeroIndex Nothing Nothing 0  -- foldlStencil is centred at 0

idStencil :: Index ix => StencilB200 ix e
idStencil = foldStencilB200 KId A.oneSz

sumStencil, productStencil :: (Num e, Index ix) => Sz ix -> StencilB200 ix e
sumStencil = foldStencilB200 KSum
productStencil = foldStencilB200 KProduct

-- | @sumStencil sz / fromIntegral (totalElem sz)@ — the device divides, it does not multiply by
-- a reciprocal, so results are bit-identical to massiv's.
avgStencil :: (Fractional e, Index ix) => Sz ix -> StencilB200 ix e
avgStencil = foldStencilB200 KAvg

maxStencil, minStencil :: (Bounded e, Ord e, Index ix) => Sz ix -> StencilB200 ix e
maxStencil = foldStencilB200 KMax
minStencil = foldStencilB200 KMin

-- | stencilCenter = size - 1 - (size `quot` 2)   (Data.Massiv.Array.Stencil.Convolution 71-73)
makeConvolutionStencilFromKernel :: (Num e, Index ix) => Array S ix e -> StencilB200 ix e
makeConvolutionStencilFromKernel k =
  StencilB200 KConv sz (A.liftIndex2 (-) (A.liftIndex (subtract 1) szi) (A.liftIndex (`quot` 2) szi)) (Just k) Nothing 0
  where sz@(Sz szi) = size k

-- | stencilCenter = size `div` 2   (Data.Massiv.Array.Stencil.Convolution 114)
makeCorrelationStencilFromKernel :: (Num e, Index ix) => Array S ix e -> StencilB200 ix e
makeCorrelationStencilFromKernel k = StencilB200 KCorr sz (A.liftIndex (`div` 2) szi) (Just k) Nothing 0
  where sz@(Sz szi) = size k

-- | The rule stencil of massiv-examples/GameOfLife/app/GameOfLife.hs lines 15-36.
lifeStencil :: StencilB200 Ix2 Word8
lifeStencil = StencilB200 KLife (Sz (3 A.:. 3)) (1 A.:. 1) Nothing Nothing 0

-- | @sqrt (fmap (^2) (makeConvolutionStencilFromKernel kx) + fmap (^2) (makeConvolutionStencilFromKernel ky))@
-- as one fused pass (massiv-bench Data.Massiv.Bench.Sobel lines 39-44).
sobelMagnitude :: (Floating e, Index ix) => Array S ix e -> Array S ix e -> StencilB200 ix e
sobelMagnitude kx ky = (makeConvolutionStencilFromKernel kx) { sbKind = KMag2, sbCoeffs2 = Just ky }

-- | The rule stencils of Data.Massiv.Array.Numeric.Integral (lines 54-118), reified: size 1 in every
-- dimension but @dim@, centre 0, @dx@ carried as the single coefficient.  The device evaluates
-- them in the reference's operation order (e.g. @(dx / 2) * ((g 0 + 2 * g 1 + ..) + g n)@).
midpointStencil, trapezoidStencil, simpsonsStencil
  :: (Fractional e, ElemB200 e, Index ix) => e -> A.Dim -> Int -> StencilB200 ix e
midpointStencil dx dim k = ruleStencil KIntMidpoint dx dim k
trapezoidStencil dx dim n = ruleStencil KIntTrapezoid dx dim (n + 1)
simpsonsStencil dx dim n
  | odd n = error $ "Number of sample points for Simpson's rule stencil should be even, but received: " ++ show n
  | otherwise = ruleStencil KIntSimpson dx dim (n + 1)

ruleStencil :: (ElemB200 e, Index ix) => Kind -> e -> A.Dim -> Int -> StencilB200 ix e
ruleStencil kind dx dim len =
  StencilB200 kind (Sz (A.setDim' (A.pureIndex 1) dim len)) A.zeroIndex (Just (A.computeAs A.S (A.singleton dx))) Nothing 0

-- | @integrateWith@ (Integral.hs 121-135): @computeWithStride (Stride nsz) . mapStencil (Fill 0) (stencil dim n)@.
integrateWithB200
  :: (Fractional e, ElemB200 e, Index ix)
  => DeviceB200 -> (A.Dim -> Int -> StencilB200 ix e) -> A.Dim -> Int -> Array S ix e -> IO (Array S ix e)
integrateWithB200 dev stencil dim n =
  computeWithStrideB200 dev (A.Stride (A.setDim' (A.pureIndex 1) dim n)) (Fill 0) (stencil dim n)

-- | Promise finite inputs: zero coefficients may be skipped (bit-identical for finite data).
assumeFinite :: StencilB200 ix e -> StencilB200 ix e
assumeFinite s = s { sbFlags = sbFlags s .|. 2 }

-- ---------------------------------------------------------------------------------------------

borderCode :: Border e -> Int32
borderCode = \case { Fill _ -> 0; Wrap -> 1; Edge -> 2; Reflect -> 3; Continue -> 4 }

ixList :: Index ix => ix -> [Int64]
ixList = map fromIntegral . A.foldlIndex (flip (:)) [] >>= \f ix -> reverse (f ix)

-- | Fill an @mb_stencil_desc@ and run an action on it.
withDesc
  :: forall ix e a. (Index ix, ElemB200 e)
  => Padding ix e -> Maybe (Stride ix) -> StencilB200 ix e -> (Ptr MbDesc -> IO a) -> IO a
withDesc (Padding (Sz po) (Sz pb) border) mstride StencilB200{..} act =
  allocaBytesAligned mbDescBytes 8 $ \p -> do
    fillBytes p 0 mbDescBytes
    let rank = A.dimensions sbSize
        pokeIx off ix = pokeArray (castPtr (p `plusPtr` off) :: Ptr Int64) (ixList ix)
    pokeByteOff p 0 (fromIntegral (fromEnum sbKind) :: Int32)
    pokeByteOff p 4 (elemCode (Nothing :: Maybe e))
    pokeByteOff p 8 (fromIntegral (A.unDim rank) :: Int32)
    pokeByteOff p 12 (borderCode border)
    pokeIx 16 (A.unSz sbSize)
    pokeIx 56 sbCenter
    pokeIx 96 po
    pokeIx 136 pb
    pokeIx 176 (maybe (A.pureIndex 1) unStride mstride)
    case border of
      Fill v -> poke (castPtr (p `plusPtr` 216)) v
      _ -> pure ()
    pokeByteOff p 256 sbFlags
    let withK Nothing k = k nullPtr
        withK (Just arr) k = A.unsafeWithPtr arr (k . castPtr)   -- Storable.hs line 278: pointer stable for the call
    withK sbCoeffs $ \k1 -> withK sbCoeffs2 $ \k2 -> do
      pokeByteOff p 224 (k1 :: Ptr ())
      pokeByteOff p 232 (k2 :: Ptr ())
      act p

check :: DeviceB200 -> CInt -> IO ()
check (DeviceB200 ctx) st = when (st /= 0) $ do
  msg <- c_mb_last_error ctx >>= peekCString
  case st of
    3 -> throwIO (IndexZeroException 0)                       -- MB_ERR_INDEX_ZERO
    2 -> throwIO (B200Error 2 ("size mismatch: " ++ msg))     -- surfaced like SizeMismatchException
    _ -> throwIO (B200Error (fromIntegral st) msg)

-- | @computeWithStride stride (applyStencil padding stencil arr)@ on the device.
computeWithStrideB200'
  :: forall ix e. (Index ix, ElemB200 e)
  => DeviceB200 -> Maybe (Stride ix) -> Padding ix e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
computeWithStrideB200' dev@(DeviceB200 ctx) mstride padding st arr =
  withDesc padding mstride st $ \pd ->
    withArray (ixList (A.unSz (size arr))) $ \pin ->
      allocaArray 5 $ \pout -> do
        c_mb_out_dims pd pin pout >>= check dev
        odims <- peekArray (A.unDim (A.dimensions (size arr))) pout
        let osz = Sz (A.foldlIndex' odims)            -- rebuild an ix from the list (helper elided)
        marr <- A.unsafeNew osz                        -- pinned: mallocPlainForeignPtrAlignedBytes
        A.unsafeWithPtr arr $ \src ->
          A.withPtr marr $ \dst ->
            c_mb_run ctx pd pin (castPtr src) (castPtr dst) >>= check dev
        A.unsafeFreeze (A.getComp arr) marr

-- | @compute (applyStencil padding stencil arr)@
applyStencilB200 :: (Index ix, ElemB200 e) => DeviceB200 -> Padding ix e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
applyStencilB200 dev = computeWithStrideB200' dev Nothing

-- | @compute (mapStencil border stencil arr)@ — samePadding: lo = centre, hi = size - (centre + 1)
mapStencilB200 :: (Index ix, ElemB200 e) => DeviceB200 -> Border e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
mapStencilB200 dev border st =
  applyStencilB200 dev (Padding (Sz (sbCenter st)) (Sz (A.liftIndex2 (-) (A.unSz (sbSize st)) (A.liftIndex (+ 1) (sbCenter st)))) border) st

computeB200 :: (Index ix, ElemB200 e) => DeviceB200 -> Border e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
computeB200 = mapStencilB200

computeWithStrideB200
  :: (Index ix, ElemB200 e) => DeviceB200 -> Stride ix -> Border e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
computeWithStrideB200 dev stride border st =
  computeWithStrideB200' dev (Just stride)
    (Padding (Sz (sbCenter st)) (Sz (A.liftIndex2 (-) (A.unSz (sbSize st)) (A.liftIndex (+ 1) (sbCenter st)))) border) st

-- | @iterateUntil (\\n _ _ -> n >= steps - 1) (\\_ -> mapStencil border stencil)@: the array stays on
-- the device between the steps, so the PCIe copies are paid once, not per step.
iterateStencilB200
  :: forall ix e. (Index ix, ElemB200 e)
  => DeviceB200 -> Int -> Border e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
iterateStencilB200 dev@(DeviceB200 ctx) steps border st arr =
  withDesc (Padding (Sz (sbCenter st)) (Sz (A.liftIndex2 (-) (A.unSz (sbSize st)) (A.liftIndex (+ 1) (sbCenter st)))) border) Nothing st $ \pd ->
    withArray (ixList (A.unSz (size arr))) $ \pin -> do
      marr <- A.unsafeNew (size arr)
      A.unsafeWithPtr arr $ \src ->
        A.withPtr marr $ \dst ->
          c_mb_iterate ctx pd pin (castPtr src) (castPtr dst) (fromIntegral steps) >>= check dev
      A.unsafeFreeze (A.getComp arr) marr
