"""poseidon252_b200 -- B200-native batched Poseidon/Hades engine with the dusk_poseidon API.

Public surface mirrors /root/reference/src/lib.rs:13-31:
    Hash, Domain, Error, HADES_WIDTH, encrypt, decrypt
plus the batch entry points this engine adds:
    Hash.digest_batch, hades.permute_batch, encrypt_batch, decrypt_batch, merkle4_build.
All computation runs in hand-written sm_100a CUDA behind the C ABI in include/poseidon252_b200.h.
"""
from . import hades, merkle, scalar
from .encryption import decrypt, decrypt_batch, encrypt, encrypt_batch
from .engine import Engine, default_engine
from .errors import (DecryptionFailed, EncryptionFailed, EngineError, Error, InvalidIOPattern, InvalidPoint,
                     IOPatternViolation, TooFewInputElements)
from .hash import Domain, Hash
from .merkle import merkle4_build, merkle4_level

HADES_WIDTH = hades.WIDTH

__all__ = ["Hash", "Domain", "Error", "HADES_WIDTH", "encrypt", "decrypt", "encrypt_batch", "decrypt_batch",
           "hades", "merkle", "scalar", "Engine", "default_engine", "merkle4_build", "merkle4_level",
           "IOPatternViolation", "InvalidIOPattern", "TooFewInputElements", "EncryptionFailed",
           "DecryptionFailed", "InvalidPoint", "EngineError"]
