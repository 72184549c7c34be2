"""Wire formats and packet-type enums of the six reference servers, as numpy structured dtypes.

Each dtype is the reference's `#pragma pack(1) struct message`, byte for byte:
  lock_2pl/udp/net.h:11-31, lock_fasst/udp/net.h:11-31, log_server/udp/net.h:15-30,
  store/udp/net.h:15-41, tatp/udp/net.h:15-65, smallbank/udp/net.h:15-52.
"""
import numpy as np

LOCK2PL, FASST, LOG, STORE, TATP, SMALLBANK = range(6)
KIND_NAMES = ["lock_2pl", "lock_fasst", "log_server", "store", "tatp", "smallbank"]

MSG_DTYPE = [
    np.dtype([("action", "u1"), ("lid", "<u4"), ("type", "u1")]),
    np.dtype([("type", "u1"), ("lid", "<u4"), ("ver", "<u4")]),
    np.dtype([("type", "u1"), ("key", "<u8"), ("val", "u1", (40,)), ("ver", "<u4")]),
    np.dtype([("type", "u1"), ("key", "<u8"), ("val", "u1", (40,)), ("ver", "<u4")]),
    np.dtype([("ord", "u1"), ("type", "u1"), ("table", "u1"), ("key", "<u8"), ("val", "u1", (40,)), ("ver", "<u4")]),
    np.dtype([("ord", "u1"), ("type", "u1"), ("table", "u1"), ("key", "<u8"), ("val", "u1", (8,)), ("ver", "<u4")]),
]
MSG_SIZE = [d.itemsize for d in MSG_DTYPE]
assert MSG_SIZE == [6, 9, 53, 53, 55, 23]

# reference struct log_entry layouts (NOT packed): log_server/udp/utils.h:19-23, tatp/udp/kvs.h:23-29,
# smallbank/udp/kvs.h:20-25
LOG_ENTRY_SIZE = [0, 0, 56, 0, 64, 32]


class Lock2pl:            # lock_2pl/udp/net.h:11-23
    kAcquireLock, kReleaseLock, kGrantLock, kRejectLock, kRetry, kReleaseAck = range(6)
    kShared, kExclusive = 0, 1


class Fasst:              # lock_fasst/udp/net.h:11-21
    kRead, kAcquireLock, kAbort, kCommit, kGrantRead, kGrantLock, kRejectLock, kAbortAck, kCommitAck = range(9)


class Log:                # log_server/udp/net.h:15-18
    kCommit, kAck = 0, 1


class Store:              # store/udp/net.h:15-29
    kRead, kSet, kInsert, kGrantRead, kRejectRead, kSetAck, kRejectSet, kNotExist, kInsertAck, kRejectInsert = range(10)


class Tatp:               # tatp/udp/net.h:15-52
    (kRead, kAcquireLock, kAbort, kCommit, kGrantRead, kRejectRead, kNotExist, kGrantLock, kRejectLock,
     kAbortAck, kCommitAck, kRejectCommit, kCommitPrim, kCommitBck, kCommitLog, kCommitPrimAck,
     kCommitBckAck, kCommitLogAck, kInsertPrim, kInsertBck, kInsertPrimAck, kInsertBckAck, kDeletePrim,
     kDeleteBck, kDeleteLog, kDeletePrimAck, kDeleteBckAck, kDeleteLogAck) = range(28)
    kSubscriber, kSecondSubscriber, kAccessInfo, kSpecialFacility, kCallForwarding = range(5)


class Smallbank:          # smallbank/udp/net.h:15-38
    (kAcquireShared, kAcquireExclusive, kReleaseShared, kReleaseExclusive, kCommitPrim, kCommitBck,
     kCommitLog, kGrantShared, kRejectShared, kGrantExclusive, kRejectExclusive, kReleaseSharedAck,
     kReleaseExclusiveAck, kCommitPrimAck, kCommitBckAck, kCommitLogAck, kRetry, kWarmupRead,
     kWarmupReadAck) = range(19)
    kSaving, kChecking = 0, 1


def as_records(kind, raw):
    """View a uint8 buffer of n*msg bytes as the structured wire dtype."""
    a = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1)
    return a.view(MSG_DTYPE[kind])


def as_bytes(rec):
    """View structured wire records as a flat uint8 array."""
    return np.ascontiguousarray(rec).view(np.uint8).reshape(-1)
