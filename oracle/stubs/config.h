/* oracle/stubs/config.h -- TEST INFRASTRUCTURE ONLY.
 * Empty stand-in for the autoconf-generated config.h that the reference
 * sources include (kernel/ipfft.h:31). Nothing from autoconf is needed to
 * compile the integer-only part of the reference. */
