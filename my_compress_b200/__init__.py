"""my_compress_b200 -- B200-native drop-in for the lz77 path of YuBinRen/my_compress."""
