// Error enums with the reference's names and cases (Sources/Deflate/DeflateError.swift:10-19, BZip2/BZip2Error.swift:12-44,
// LZMA/LZMAError.swift:10-25, LZMA2/LZMA2Error.swift:10-22, Common/DataError.swift:9-25, GZip/GzipError.swift:10-35,
// Zlib/ZlibError.swift:12-26, XZ/XZError.swift:12-48).  A status code of include/swc_status.h is <enum base> + 1-based case index.
import Foundation

public enum DeflateError: Error { case wrongUncompressedBlockLengths, wrongBlockType, wrongSymbol, symbolNotFound }
public enum BZip2Error: Error {
    case wrongMagic, wrongVersion, wrongBlockSize, wrongBlockType, randomizedBlock, wrongHuffmanGroups, wrongSelector,
         wrongHuffmanCodeLength, symbolNotFound
    case wrongCRC(Data)
}
public enum LZMAError: Error {
    case wrongProperties, rangeDecoderInitError, exceededUncompressedSize, windowIsEmpty, rangeDecoderFinishError,
         repeatWillExceed, notEnoughToRepeat
}
public enum LZMA2Error: Error { case wrongDictionarySize, wrongControlByte, wrongReset, wrongSizes }
public enum DataError: Error, Equatable { case truncated, corrupted, checksumMismatch([Data]), unsupportedFeature }
public enum GzipError: Error {
    case wrongMagic, wrongCompressionMethod, wrongFlags, wrongHeaderCRC
    case wrongCRC([Data])
    case wrongISize, cannotEncodeISOLatin1
}
public enum ZlibError: Error {
    case wrongCompressionMethod, wrongCompressionInfo, wrongFcheck, wrongCompressionLevel
    case wrongAdler32(Data)
}
public enum XZError: Error {
    case wrongMagic, wrongField, wrongInfoCRC, wrongFilterID, checkTypeSHA256, wrongDataSize
    case wrongCheck([Data])
    case wrongPadding, multiByteIntegerError
}
/// Engine conditions that have no case in the reference (no device, CUDA failure, input on which the reference traps).
public enum SWCGPUError: Error { case outputOverflow, referenceTrap, cuda(String), invalidArgument, noDevice, unsupported }

func swcError(_ status: Int32, payload: [Data] = []) -> Error {
    let p = payload.first ?? Data()
    switch status {
    case 1: return SWCGPUError.outputOverflow
    case 2: return SWCGPUError.referenceTrap
    case 3: return SWCGPUError.cuda(String(cString: swc_last_error_string()))
    case 4: return SWCGPUError.invalidArgument
    case 5: return SWCGPUError.noDevice
    case 101: return DeflateError.wrongUncompressedBlockLengths
    case 102: return DeflateError.wrongBlockType
    case 103: return DeflateError.wrongSymbol
    case 104: return DeflateError.symbolNotFound
    case 201: return BZip2Error.wrongMagic
    case 202: return BZip2Error.wrongVersion
    case 203: return BZip2Error.wrongBlockSize
    case 204: return BZip2Error.wrongBlockType
    case 205: return BZip2Error.randomizedBlock
    case 206: return BZip2Error.wrongHuffmanGroups
    case 207: return BZip2Error.wrongSelector
    case 208: return BZip2Error.wrongHuffmanCodeLength
    case 209: return BZip2Error.symbolNotFound
    case 210: return BZip2Error.wrongCRC(p)
    case 301: return LZMAError.wrongProperties
    case 302: return LZMAError.rangeDecoderInitError
    case 303: return LZMAError.exceededUncompressedSize
    case 304: return LZMAError.windowIsEmpty
    case 305: return LZMAError.rangeDecoderFinishError
    case 306: return LZMAError.repeatWillExceed
    case 307: return LZMAError.notEnoughToRepeat
    case 401: return LZMA2Error.wrongDictionarySize
    case 402: return LZMA2Error.wrongControlByte
    case 403: return LZMA2Error.wrongReset
    case 404: return LZMA2Error.wrongSizes
    case 501: return DataError.truncated
    case 502: return DataError.corrupted
    case 503: return DataError.checksumMismatch(payload)
    case 504: return DataError.unsupportedFeature
    case 601: return GzipError.wrongMagic
    case 602: return GzipError.wrongCompressionMethod
    case 603: return GzipError.wrongFlags
    case 604: return GzipError.wrongHeaderCRC
    case 605: return GzipError.wrongCRC(payload)
    case 606: return GzipError.wrongISize
    case 701: return ZlibError.wrongCompressionMethod
    case 702: return ZlibError.wrongCompressionInfo
    case 703: return ZlibError.wrongFcheck
    case 704: return ZlibError.wrongCompressionLevel
    case 705: return ZlibError.wrongAdler32(p)
    case 801: return XZError.wrongMagic
    case 802: return XZError.wrongField
    case 803: return XZError.wrongInfoCRC
    case 804: return XZError.wrongFilterID
    case 805: return XZError.checkTypeSHA256
    case 806: return XZError.wrongDataSize
    case 807: return XZError.wrongCheck(payload)
    case 808: return XZError.wrongPadding
    case 809: return XZError.multiByteIntegerError
    default: return SWCGPUError.unsupported
    }
}
