"""is the hardware video decoder reachable on this box?  dlopen of the driver's
libnvcuvid / libnvidia-encode with a hand-declared subset of the Video Codec API (no SDK headers in the image):
cuvidGetDecoderCaps for H.264 / HEVC 8-bit 4:2:0.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class CUVIDDECODECAPS(C.Structure):
    _fields_ = [("eCodecType", C.c_int), ("eChromaFormat", C.c_int), ("nBitDepthMinus8", C.c_uint), ("reserved1", C.c_uint * 3),
                ("bIsSupported", C.c_ubyte), ("nNumNVDECs", C.c_ubyte), ("nOutputFormatMask", C.c_ushort), ("nMaxWidth", C.c_uint),
                ("nMaxHeight", C.c_uint), ("nMaxMBCount", C.c_uint), ("nMinWidth", C.c_ushort), ("nMinHeight", C.c_ushort),
                ("bIsHistogramSupported", C.c_ubyte), ("nCounterBitDepth", C.c_ubyte), ("nMaxHistogramBins", C.c_ushort),
                ("reserved3", C.c_uint * 10)]


def main():
    out = {"libnvcuvid": None, "libnvidia-encode": None, "caps": {}}
    for key, names in (("libnvcuvid", ("libnvcuvid.so.1", "libnvcuvid.so")), ("libnvidia-encode", ("libnvidia-encode.so.1", "libnvidia-encode.so"))):
        for nm in names:
            try:
                C.CDLL(nm)
                out[key] = nm
                break
            except OSError as e:
                out[key + "_error"] = str(e)[:120]
    if out["libnvcuvid"]:
        try:
            import torch
            torch.zeros(1, device="cuda")                      # primary context current on this thread
            lib = C.CDLL(out["libnvcuvid"])
            lib.cuvidGetDecoderCaps.restype = C.c_int
            for codec, cid in (("h264", 4), ("hevc", 8), ("av1", 11)):
                caps = CUVIDDECODECAPS()
                caps.eCodecType, caps.eChromaFormat, caps.nBitDepthMinus8 = cid, 1, 0      # cudaVideoChromaFormat_420
                rc = lib.cuvidGetDecoderCaps(C.byref(caps))
                out["caps"][codec] = {"rc": rc, "supported": int(caps.bIsSupported), "engines": int(caps.nNumNVDECs),
                                      "max": [int(caps.nMaxWidth), int(caps.nMaxHeight)]}
        except Exception as e:                                   # noqa: BLE001
            out["caps_error"] = repr(e)[:200]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
