{-# LANGUAGE BangPatterns #-}
{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE LambdaCase #-}
{-# LANGUAGE MagicHash #-}
{-# LANGUAGE RankNTypes #-}
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
-- replaces @compute@ of the @Array DW@ that 'mapStencil' would have produced.  'StencilB200' has the
-- same 'Functor'-like / 'Num' / 'Fractional' / 'Floating' algebra as massiv's 'Stencil'
-- (Internal.hs lines 93-174) for the operations whose result IEEE-754 / two's complement fixes:
-- @sX * sX + sY * sY@, @sqrt@, @abs (a - b)@, @s / 9@ ... build an expression that the device evaluates in
-- one pass (MB_EXPR).  Closure stencils ('makeStencil') and libm functions (@exp@, @sin@ ...) have no
-- device form: they simply are not constructible here, there is no silent CPU fallback.
--
-- NOTE: this file is shipped as source only.  The build image has no GHC (SURVEY.md, fact 2), so it
-- has not been compiled; INTEGRATION.md lists the steps a maintainer runs.  The C ABI it binds is
-- exercised by tests/ (through ctypes, through a plain C harness, tests/cabi_harness.c, and through the C++
-- mirror include/massiv_b200.hpp); the structure offsets used below are the ones tests/test_abi_layout.py
-- checks against @offsetof@.
module Data.Massiv.Array.Stencil.B200
  ( -- * Device
    DeviceB200
  , withDeviceB200
  , trimDeviceB200
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
  , makeConvolutionStencil
  , makeCorrelationStencil
  , lifeStencil
  , minStencils
  , maxStencils
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
  , computeSeparableB200
  , iterateStencilB200
  , UntilB200(..)
  , iterateUntilB200
    -- * Primitive arrays
  , mapStencilB200P
    -- * Sharded iteration (one process per GPU)
  , ncclUniqueIdB200
  , shardInitB200
  , iterateShardedB200
    -- * Errors
  , B200Exception(..)
  ) where

import Control.Exception (Exception, bracket, throwIO)
import Control.Monad (forM_, when, zipWithM_)
import Data.Massiv.Array
  ( Array, Border(..), Ix2, Index, P, Padding(..), S, Stride, Sz(..), size, unStride )
import qualified Data.Massiv.Array as A
import qualified Data.Massiv.Array.Unsafe as A
import Data.Massiv.Core.Index (IndexException(..))
import qualified Data.ByteString as BS
import qualified Data.ByteString.Unsafe as BS
import Data.Int
import Data.List (elemIndex)
import qualified Data.Primitive.ByteArray as BA
import Data.Word
import Foreign
import Foreign.C.String (CString, peekCString)
import Foreign.C.Types

-- ---------------------------------------------------------------------------------------------
-- C ABI (include/massiv_b200.h).  All calls are `safe`: they last milliseconds to seconds and an
-- `unsafe` call would stall every capability and block GC.
-- ---------------------------------------------------------------------------------------------

data MbCtx
data MbDesc

foreign import ccall safe "mb_ctx_create"  c_mb_ctx_create  :: CInt -> Ptr (Ptr MbCtx) -> IO CInt
foreign import ccall safe "mb_ctx_destroy" c_mb_ctx_destroy :: Ptr MbCtx -> IO CInt
-- | gives the context's device scratch (staging arrays, padded iteration pairs) back to the driver
foreign import ccall safe "mb_ctx_trim"    c_mb_ctx_trim    :: Ptr MbCtx -> IO CInt
foreign import ccall safe "mb_last_error"  c_mb_last_error  :: Ptr MbCtx -> IO CString
foreign import ccall safe "mb_out_dims"    c_mb_out_dims    :: Ptr MbDesc -> Ptr Int64 -> Ptr Int64 -> IO CInt
foreign import ccall safe "mb_run"         c_mb_run
  :: Ptr MbCtx -> Ptr MbDesc -> Ptr Int64 -> Ptr () -> Ptr () -> IO CInt
foreign import ccall safe "mb_run_sep2"    c_mb_run_sep2
  :: Ptr MbCtx -> Ptr MbDesc -> Ptr MbDesc -> Ptr Int64 -> Ptr () -> Ptr () -> IO CInt
foreign import ccall safe "mb_iterate"     c_mb_iterate
  :: Ptr MbCtx -> Ptr MbDesc -> Ptr Int64 -> Ptr () -> Ptr () -> Int64 -> IO CInt
foreign import ccall safe "mb_iterate_until" c_mb_iterate_until
  :: Ptr MbCtx -> Ptr MbDesc -> Ptr Int64 -> Ptr () -> Ptr () -> Int64 -> Int64 -> CInt -> CDouble
  -> Ptr Int64 -> Ptr Int32 -> IO CInt
foreign import ccall safe "mb_nccl_unique_id" c_mb_nccl_unique_id :: Ptr Word8 -> IO CInt
foreign import ccall safe "mb_shard_init"  c_mb_shard_init  :: Ptr MbCtx -> CInt -> CInt -> Ptr Word8 -> IO CInt
foreign import ccall safe "mb_iterate_sharded" c_mb_iterate_sharded
  :: Ptr MbCtx -> Ptr MbDesc -> Ptr Int64 -> Ptr Int64 -> Ptr () -> Ptr () -> Int64 -> IO CInt

-- | sizes of @mb_stencil_desc@, @mb_term@, @mb_xins@, @mb_post_op@ (INTEGRATION.md section 2)
mbDescBytes, mbTermBytes, mbXinsBytes :: Int
mbDescBytes = 296
mbTermBytes = 112
mbXinsBytes = 16

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

-- | Returns the context's grow-only device scratch (staging arrays of the host-to-host calls, the padded pair of an
-- iteration over a volume TMA cannot describe) to the driver.
trimDeviceB200 :: DeviceB200 -> IO ()
trimDeviceB200 (DeviceB200 p) = do
  st <- c_mb_ctx_trim p
  when (st /= 0) $ throwIO (B200Error (fromIntegral st) "mb_ctx_trim")

-- | @mb_kind@
data Kind = KId | KSum | KProduct | KAvg | KMax | KMin | KConv | KCorr | KTaps | KLife | KMag2
          | KIntMidpoint | KIntTrapezoid | KIntSimpson | KExpr
  deriving (Eq, Enum)

-- | @mb_xop@ (program steps of an expression; the numbering starts at 1)
data XOp = XTerm | XConst | XAdd | XSub | XMul | XDiv | XMin | XMax | XNegate | XAbs | XSignum | XSquare | XRecip | XSqrt
  deriving (Eq, Enum)

xopCode :: XOp -> Int32
xopCode = (+ 1) . fromIntegral . fromEnum

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

-- | One built-in constructor, reified (a "term" of an expression).
data Base ix e = Base
  { bKind   :: !Kind
  , bSize   :: !(Sz ix)
  , bCenter :: !ix
  , bCoeffs :: Maybe (Array S ix e)       -- FromKernel forms: the kernel array, row-major
  , bTaps   :: Maybe ([ix], [e])          -- closure forms: effective source offsets and coefficients, application order
  }

-- | The reified counterpart of @Stencil ix e e@: an expression over built-in constructors.
data StencilB200 ix e
  = Term !(Base ix e)
  | Pure !e                                         -- pure / fromInteger / fromRational (Internal.hs 94-95, 127, 135)
  | Bin !XOp (StencilB200 ix e) (StencilB200 ix e)  -- liftA2 (Internal.hs 104-111)
  | Un !XOp (StencilB200 ix e)                      -- fmap (Internal.hs 36-41)
  | Finite (StencilB200 ix e)                       -- the caller's promise of finite inputs (MB_FLAG_ASSUME_FINITE)

-- | The instances of Internal.hs lines 114-146, for the methods whose results are exactly specified.
instance (Index ix, Num e) => Num (StencilB200 ix e) where
  (+) = Bin XAdd
  (-) = Bin XSub
  (*) = Bin XMul
  negate = Un XNegate
  abs = Un XAbs
  signum = Un XSignum
  fromInteger = Pure . fromInteger

instance (Index ix, Fractional e) => Fractional (StencilB200 ix e) where
  (/) = Bin XDiv
  recip = Un XRecip
  fromRational = Pure . fromRational

-- | Only @sqrt@ lowers: the other 'Floating' methods are libm-defined and cannot be bit-identical.
instance (Index ix, Floating e) => Floating (StencilB200 ix e) where
  sqrt = Un XSqrt
  pi = Pure pi
  exp = noDevice "exp"; log = noDevice "log"; sin = noDevice "sin"; cos = noDevice "cos"
  asin = noDevice "asin"; acos = noDevice "acos"; atan = noDevice "atan"
  sinh = noDevice "sinh"; cosh = noDevice "cosh"; asinh = noDevice "asinh"; acosh = noDevice "acosh"; atanh = noDevice "atanh"

noDevice :: String -> a
noDevice f = error ("Data.Massiv.Array.Stencil.B200: " ++ f ++ " is defined by libm and has no bit-identical device form")

-- | @liftA2 min@ / @liftA2 max@ (GHC.Classes' defaults: @min x y = if x <= y then x else y@)
minStencils, maxStencils :: StencilB200 ix e -> StencilB200 ix e -> StencilB200 ix e
minStencils = Bin XMin
maxStencils = Bin XMax

foldStencilB200 :: Index ix => Kind -> Sz ix -> StencilB200 ix e
foldStencilB200 k sz = Term (Base k sz A.zeroIndex Nothing Nothing)  -- foldlStencil is centred at 0

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
  Term (Base KConv sz (A.liftIndex2 (-) (A.liftIndex (subtract 1) szi) (A.liftIndex (`quot` 2) szi)) (Just k) Nothing)
  where sz@(Sz szi) = size k

-- | stencilCenter = size `div` 2   (Data.Massiv.Array.Stencil.Convolution 114)
makeCorrelationStencilFromKernel :: (Num e, Index ix) => Array S ix e -> StencilB200 ix e
makeCorrelationStencilFromKernel k = Term (Base KCorr sz (A.liftIndex (`div` 2) szi) (Just k) Nothing)
  where sz@(Sz szi) = size k

-- | The closure forms (Convolution.hs 44-57, 89-100).  massiv's argument has type
-- @(ix -> e -> e -> e) -> e -> e@ and is written @\\f -> f o1 k1 . f o2 k2 ...@; the same expression is
-- polymorphic in the accumulator, which is how the tap list is read off here.  @f a . f b $ 0@ applies b first:
-- the list is in application order.  A convolution takes its values at @ix - o@ and inverts the centre.
makeConvolutionStencil, makeCorrelationStencil
  :: (Num e, Index ix) => Sz ix -> ix -> (forall a. (ix -> e -> a -> a) -> a -> a) -> StencilB200 ix e
makeConvolutionStencil sz@(Sz szi) c rel =
  Term (Base KTaps sz (A.liftIndex2 (-) (A.liftIndex (subtract 1) szi) c) Nothing (Just (map (A.liftIndex negate) offs, ks)))
  where (offs, ks) = unzip (rel (\o k acc -> acc ++ [(o, k)]) [])
makeCorrelationStencil sz c rel = Term (Base KTaps sz c Nothing (Just (unzip (rel (\o k acc -> acc ++ [(o, k)]) []))))

-- | The rule stencil of massiv-examples/GameOfLife/app/GameOfLife.hs lines 15-36.
lifeStencil :: StencilB200 Ix2 Word8
lifeStencil = Term (Base KLife (Sz (3 A.:. 3)) (1 A.:. 1) Nothing Nothing)

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
  Term (Base kind (Sz (A.setDim' (A.pureIndex 1) dim len)) A.zeroIndex (Just (A.computeAs A.S (A.singleton dx))) Nothing)

-- | @integrateWith@ (Integral.hs 121-135): @computeWithStride (Stride nsz) . mapStencil (Fill 0) (stencil dim n)@.
integrateWithB200
  :: (Fractional e, ElemB200 e, Index ix)
  => DeviceB200 -> (A.Dim -> Int -> StencilB200 ix e) -> A.Dim -> Int -> Array S ix e -> IO (Array S ix e)
integrateWithB200 dev stencil dim n =
  computeWithStrideB200 dev (A.Stride (A.setDim' (A.pureIndex 1) dim n)) (Fill 0) (stencil dim n)

-- | Promise finite inputs: zero coefficients may be skipped (bit-identical for finite data).  Never needed for
-- correctness or for the speed of the 3-D star kernels (they check on the device, DESIGN.md section 3).
assumeFinite :: StencilB200 ix e -> StencilB200 ix e
assumeFinite = Finite

-- ---------------------------------------------------------------------------------------------
-- geometry: unionStencilCenters / unionStencilSizes (Internal.hs 83-90)

stencilCenterB200 :: Index ix => StencilB200 ix e -> ix
stencilCenterB200 = \case
  Term b -> bCenter b
  Pure _ -> A.zeroIndex
  Bin _ a b -> A.liftIndex2 max (stencilCenterB200 a) (stencilCenterB200 b)
  Un _ a -> stencilCenterB200 a
  Finite a -> stencilCenterB200 a

stencilSizeB200 :: Index ix => StencilB200 ix e -> Sz ix
stencilSizeB200 st = case st of
  Term b | bKind b == KAvg -> Sz (A.liftIndex (max 1) (A.unSz (bSize b)))   -- liftA2 (/) with `pure n`
         | otherwise -> bSize b
  Pure _ -> A.oneSz
  Bin _ a b ->
    let mc = stencilCenterB200 st
        ext s = A.liftIndex2 (-) (A.unSz (stencilSizeB200 s)) (stencilCenterB200 s)
    in Sz (A.liftIndex2 (+) mc (A.liftIndex2 max (ext a) (ext b)))
  Un _ a -> stencilSizeB200 a
  Finite a -> stencilSizeB200 a

samePaddingB200 :: Index ix => StencilB200 ix e -> Border e -> Padding ix e
samePaddingB200 st = Padding (Sz c) (Sz (A.liftIndex2 (-) (A.unSz (stencilSizeB200 st)) (A.liftIndex (+ 1) c)))
  where c = stencilCenterB200 st

-- ---------------------------------------------------------------------------------------------
-- lowering

borderCode :: Border e -> Int32
borderCode = \case { Fill _ -> 0; Wrap -> 1; Edge -> 2; Reflect -> 3; Continue -> 4 }

-- | outermost dimension first, the library's convention
ixList :: Index ix => ix -> [Int64]
ixList = map fromIntegral . A.foldrIndex (:) []

-- | distinct terms (by position in the expression) and the postfix program
flatten :: StencilB200 ix e -> ([Base ix e], [(XOp, Int, Maybe e)], Word32)
flatten st0 = (reverse ts, reverse prog, fl)
  where
    (ts, prog, fl) = go st0 ([], [], 0)
    go s (terms, p, f) = case s of
      Term b -> (b : terms, (XTerm, length terms, Nothing) : p, f)
      Pure c -> (terms, (XConst, 0, Just c) : p, f)
      Bin op a b -> let (t1, p1, f1) = go a (terms, p, f); (t2, p2, f2) = go b (t1, p1, f1) in (t2, (op, 0, Nothing) : p2, f2)
      Un op a -> let (t1, p1, f1) = go a (terms, p, f) in (t1, (op, 0, Nothing) : p1, f1)
      Finite a -> let (t1, p1, f1) = go a (terms, p, f) in (t1, p1, f1 .|. 2)

withCoeffs :: forall ix e a. (Index ix, ElemB200 e) => Base ix e -> (Ptr () -> Int64 -> Ptr Int64 -> IO a) -> IO a
withCoeffs Base{..} k = case (bCoeffs, bTaps) of
  (Just arr, _) -> A.unsafeWithPtr arr $ \p -> k (castPtr p) 0 nullPtr     -- Storable.hs line 278: stable for the call
  (_, Just (offs, ks)) ->
    withArray ks $ \pk -> withArray (concatMap ixList offs) $ \po -> k (castPtr pk) (fromIntegral (length ks)) po
  _ -> k nullPtr 0 nullPtr

-- | Fill an @mb_stencil_desc@ (offsets: INTEGRATION.md section 2) and run an action on it.
withDesc
  :: forall ix e a. (Index ix, ElemB200 e)
  => Padding ix e -> Maybe (Stride ix) -> StencilB200 ix e -> (Ptr MbDesc -> IO a) -> IO a
withDesc (Padding (Sz po) (Sz pb) border) mstride st act =
  allocaBytesAligned mbDescBytes 8 $ \p -> do
    fillBytes p 0 mbDescBytes
    let pokeIx :: Ptr b -> Int -> ix -> IO ()
        pokeIx base off ix = pokeArray (castPtr (base `plusPtr` off) :: Ptr Int64) (ixList ix)
        common kind sz c = do
          pokeByteOff p 0 (fromIntegral (fromEnum kind) :: Int32)
          pokeByteOff p 4 (elemCode (Nothing :: Maybe e))
          pokeByteOff p 8 (fromIntegral (A.unDim (A.dimensions sz)) :: Int32)
          pokeByteOff p 12 (borderCode border)
          pokeIx p 16 (A.unSz sz)
          pokeIx p 56 c
          pokeIx p 96 po
          pokeIx p 136 pb
          pokeIx p 176 (maybe (A.pureIndex 1) unStride mstride)
          case border of
            Fill v -> poke (castPtr (p `plusPtr` 216)) v
            _ -> pure ()
    case flatten st of
      -- a bare constructor keeps its own kind (and the kernels specialised for it)
      ([b], [(XTerm, 0, Nothing)], fl) -> do
        common (bKind b) (bSize b) (bCenter b)
        pokeByteOff p 256 fl
        withCoeffs b $ \k1 ntaps offs -> do
          pokeByteOff p 224 k1
          pokeByteOff p 240 ntaps
          pokeByteOff p 248 offs
          act p
      (terms, prog, fl) -> do
        common KExpr (stencilSizeB200 st) (stencilCenterB200 st)
        pokeByteOff p 256 fl
        allocaBytesAligned (mbTermBytes * length terms) 8 $ \pt ->
          allocaBytesAligned (mbXinsBytes * length prog) 8 $ \px -> do
            fillBytes pt 0 (mbTermBytes * length terms)
            fillBytes px 0 (mbXinsBytes * length prog)
            forM_ (zip [0 ..] prog) $ \(i, (op, arg, mc)) -> do
              let q = px `plusPtr` (i * mbXinsBytes)
              pokeByteOff q 0 (xopCode op)
              pokeByteOff q 4 (fromIntegral arg :: Int32)
              maybe (pure ()) (poke (castPtr (q `plusPtr` 8))) mc
            pokeByteOff p 272 (fromIntegral (length terms) :: Word32)
            pokeByteOff p 276 (fromIntegral (length prog) :: Word32)
            pokeByteOff p 280 pt
            pokeByteOff p 288 px
            -- every term's kernel / tap list stays pinned for the duration of the call
            let fillTerm i b rest = withCoeffs b $ \k1 ntaps offs -> do
                  let q = pt `plusPtr` (i * mbTermBytes)
                  pokeByteOff q 0 (fromIntegral (fromEnum (bKind b)) :: Int32)
                  pokeIx q 8 (A.unSz (bSize b))
                  pokeIx q 48 (bCenter b)
                  pokeByteOff q 88 k1
                  pokeByteOff q 96 ntaps
                  pokeByteOff q 104 offs
                  rest
            foldr (\(i, b) rest -> fillTerm i b rest) (act p) (zip [0 ..] terms)

check :: DeviceB200 -> CInt -> IO ()
check (DeviceB200 ctx) st = when (st /= 0) $ do
  msg <- c_mb_last_error ctx >>= peekCString
  case st of
    3 -> throwIO (IndexZeroException 0)                       -- MB_ERR_INDEX_ZERO
    2 -> throwIO (B200Error 2 ("size mismatch: " ++ msg))     -- surfaced like SizeMismatchException
    _ -> throwIO (B200Error (fromIntegral st) msg)

listToIx :: forall ix. Index ix => [Int64] -> ix
listToIx ds = snd (A.iterLinear' ds)   -- rebuild an ix from outermost-first extents (helper: zip with setDim')

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
        marr <- A.unsafeNew (Sz (listToIx odims))      -- pinned: mallocPlainForeignPtrAlignedBytes
        A.unsafeWithPtr arr $ \src ->
          A.withPtr marr $ \dst ->
            c_mb_run ctx pd pin (castPtr src) (castPtr dst) >>= check dev
        A.unsafeFreeze (A.getComp arr) marr

-- | @compute (applyStencil padding stencil arr)@
applyStencilB200 :: (Index ix, ElemB200 e) => DeviceB200 -> Padding ix e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
applyStencilB200 dev = computeWithStrideB200' dev Nothing

-- | @compute (mapStencil border stencil arr)@ — samePadding: lo = centre, hi = size - (centre + 1)
mapStencilB200 :: (Index ix, ElemB200 e) => DeviceB200 -> Border e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
mapStencilB200 dev border st = applyStencilB200 dev (samePaddingB200 st border) st

computeB200 :: (Index ix, ElemB200 e) => DeviceB200 -> Border e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
computeB200 = mapStencilB200

computeWithStrideB200
  :: (Index ix, ElemB200 e) => DeviceB200 -> Stride ix -> Border e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
computeWithStrideB200 dev stride border st = computeWithStrideB200' dev (Just stride) (samePaddingB200 st border) st

-- | @compute . mapStencil b second . compute . mapStencil b first@: the two-pass idiom for rank-1 factorable
-- kernels, one pass over HBM on the device (the intermediate is rounded to @e@ exactly as the two computes do).
computeSeparableB200
  :: (Index ix, ElemB200 e) => DeviceB200 -> Border e -> StencilB200 ix e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
computeSeparableB200 dev@(DeviceB200 ctx) border first second arr =
  withDesc (samePaddingB200 first border) Nothing first $ \p1 ->
    withDesc (samePaddingB200 second border) Nothing second $ \p2 ->
      withArray (ixList (A.unSz (size arr))) $ \pin -> do
        marr <- A.unsafeNew (size arr)
        A.unsafeWithPtr arr $ \src -> A.withPtr marr $ \dst ->
          c_mb_run_sep2 ctx p1 p2 pin (castPtr src) (castPtr dst) >>= check dev
        A.unsafeFreeze (A.getComp arr) marr

-- | @iterateUntil (\\n _ _ -> n >= steps - 1) (\\_ -> mapStencil border stencil)@: the array stays on
-- the device between the steps, so the PCIe copies are paid once, not per step.
iterateStencilB200
  :: forall ix e. (Index ix, ElemB200 e)
  => DeviceB200 -> Int -> Border e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e)
iterateStencilB200 dev@(DeviceB200 ctx) steps border st arr =
  withDesc (samePaddingB200 st border) Nothing st $ \pd ->
    withArray (ixList (A.unSz (size arr))) $ \pin -> do
      marr <- A.unsafeNew (size arr)
      A.unsafeWithPtr arr $ \src ->
        A.withPtr marr $ \dst ->
          c_mb_iterate ctx pd pin (castPtr src) (castPtr dst) (fromIntegral steps) >>= check dev
      A.unsafeFreeze (A.getComp arr) marr

-- | The convergence predicates of 'iterateUntil' (Manifest.Internal lines 331-349) the device can evaluate.
data UntilB200
  = UntilEqual              -- ^ @\\_ prev cur -> cur == prev@
  | UntilMaxAbsDiff Double  -- ^ @\\_ prev cur -> maximum (abs (cur - prev)) < tol@

-- | @iterateUntil pred (\\_ -> mapStencil border stencil)@, the predicate checked after every @checkEvery@-th step, at
-- most @maxSteps@ steps.  Returns the array, the number of steps made and whether the predicate held.
iterateUntilB200
  :: forall ix e. (Index ix, ElemB200 e)
  => DeviceB200 -> UntilB200 -> Int -> Int -> Border e -> StencilB200 ix e -> Array S ix e -> IO (Array S ix e, Int, Bool)
iterateUntilB200 dev@(DeviceB200 ctx) pr maxSteps checkEvery border st arr =
  withDesc (samePaddingB200 st border) Nothing st $ \pd ->
    withArray (ixList (A.unSz (size arr))) $ \pin -> alloca $ \psteps -> alloca $ \pconv -> do
      marr <- A.unsafeNew (size arr)
      let (code, tol) = case pr of { UntilEqual -> (0, 0); UntilMaxAbsDiff t -> (1, t) }
      A.unsafeWithPtr arr $ \src -> A.withPtr marr $ \dst ->
        c_mb_iterate_until ctx pd pin (castPtr src) (castPtr dst) (fromIntegral maxSteps) (fromIntegral checkEvery) code
                           (realToFrac tol) psteps pconv >>= check dev
      out <- A.unsafeFreeze (A.getComp arr) marr
      n <- peek psteps
      c <- peek pconv
      pure (out, fromIntegral n, c /= 0)

-- ---------------------------------------------------------------------------------------------
-- Primitive (P) arrays (Array/Manifest/Primitive.hs lines 80-88): a ByteArray and an element offset.  GHC pins byte
-- arrays larger than ~3 KiB and those made with newPinnedByteArray; a pinned one is used in place, an unpinned
-- one goes through `convert` to S (Manifest/Internal.hs lines 178-184), the one O(n) copy there is no way around.

mapStencilB200P
  :: forall ix e. (Index ix, ElemB200 e, A.Prim e)
  => DeviceB200 -> Border e -> StencilB200 ix e -> Array P ix e -> IO (Array S ix e)
mapStencilB200P dev border st parr
  | BA.isByteArrayPinned ba = do
      -- in place: the payload address is stable; unsafeArrayFromForeignPtr0 wraps it without copying
      let addr = BA.byteArrayContents ba `plusPtr` (A.unsafeLinearIndexOffset parr * sizeOf (undefined :: e))
      fp <- newForeignPtr_ (castPtr addr)
      r <- mapStencilB200 dev border st (A.unsafeArrayFromForeignPtr0 (A.getComp parr) fp (size parr))
      BA.touch ba   -- keeps the ByteArray alive until the device call has returned
      pure r
  | otherwise = mapStencilB200 dev border st (A.convert parr)
  where ba = A.toByteArray parr

-- ---------------------------------------------------------------------------------------------
-- Sharded iteration: the global array is cut along its outermost dimension, one process (and context) per GPU.
-- The 128-byte id of process 0 reaches the others by whatever means the application has (MPI, a file, a socket).

ncclUniqueIdB200 :: IO BS.ByteString
ncclUniqueIdB200 = allocaBytes 128 $ \p -> do
  st <- c_mb_nccl_unique_id p
  when (st /= 0) $ throwIO (B200Error (fromIntegral st) "mb_nccl_unique_id")
  BS.packCStringLen (castPtr p, 128)

shardInitB200 :: DeviceB200 -> Int -> Int -> BS.ByteString -> IO ()
shardInitB200 dev@(DeviceB200 ctx) nRanks rank ident =
  BS.unsafeUseAsCString ident $ \p -> c_mb_shard_init ctx (fromIntegral nRanks) (fromIntegral rank) (castPtr p) >>= check dev

-- | @steps@ applications of @mapStencil border stencil@ to the global array of extent @globalSz@ whose slab (own
-- planes only: rows @[rank * n / ranks, (rank + 1) * n / ranks)@ of the outermost dimension) this process holds.
iterateShardedB200
  :: forall ix e. (Index ix, ElemB200 e)
  => DeviceB200 -> Int -> Border e -> StencilB200 ix e -> Sz ix -> Array S ix e -> IO (Array S ix e)
iterateShardedB200 dev@(DeviceB200 ctx) steps border st (Sz globalSz) slab =
  withDesc (samePaddingB200 st border) Nothing st $ \pd ->
    withArray (ixList globalSz) $ \pglobal ->
      withArray (ixList (A.unSz (size slab))) $ \plocal -> do
        marr <- A.unsafeNew (size slab)
        A.unsafeWithPtr slab $ \src -> A.withPtr marr $ \dst ->
          c_mb_iterate_sharded ctx pd pglobal plocal (castPtr src) (castPtr dst) (fromIntegral steps) >>= check dev
        A.unsafeFreeze (A.getComp slab) marr
