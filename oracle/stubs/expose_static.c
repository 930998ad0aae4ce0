/* oracle/stubs/expose_static.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiled INSTEAD of the reference's api/api-basic.c: it includes that file from where it lies
 * under $(REF) (nothing is copied) and adds entry points that reach two of its static functions,
 * twiddle_input / twiddle_output (api/api-basic.c:1186-1285: the +-1 modulations behind
 * PFFT_SHIFTED_IN / PFFT_SHIFTED_OUT), on a hand-filled plan record, so that golden vectors of
 * the shift conventions can be captured without FFTW (tests/golden/gen_shift_golden.py). */
#include "api-basic.c"

void oracle_ref_twiddle(int output_side, int rnk_n, int rnk_pm, INT *n, INT *nio, INT *local_n, INT *local_start,
                        int *skip_trafos, INT howmany, unsigned trafo_flag, unsigned transp_flag, unsigned pfft_flags,
                        R *in, R *out)
{
  plan_s p;
  memset(&p, 0, sizeof p);
  p.rnk_n = rnk_n;
  p.rnk_pm = rnk_pm;
  p.n = n;
  p.howmany = howmany;
  p.trafo_flag = trafo_flag;
  p.transp_flag = transp_flag;
  p.pfft_flags = pfft_flags;
  p.skip_trafos = skip_trafos;
  if (output_side) {
    p.no = nio; p.local_no = local_n; p.local_no_start = local_start;
    p.otwiddle_in = in; p.otwiddle_out = out;
    twiddle_output(&p, in, out, in, out);
  } else {
    p.ni = nio; p.local_ni = local_n; p.local_ni_start = local_start;
    p.itwiddle_in = in; p.itwiddle_out = out;
    twiddle_input(&p, in, out, in, out);
  }
}
