/* empty stub: the reference includes this Crypto++ header but uses nothing from it (SEAL/bidder.h:8-12). */
