#!/usr/bin/env python
"""TEST INFRASTRUCTURE: apply the INTEGRATION.md section-1 patch to the reference's OWN diffusion_2D/main.cpp.

    python tests/native/patch_reference_main.py /root/reference/diffusion_2D/main.cpp out.cpp

The patched file is generated at build time into oracle/_ref/ (git-ignored) -- no reference source is committed.
Every substitution asserts how often its pattern occurs, so a change of the reference that moves a spot fails here
instead of silently building the wrong program.  What changes (reference line numbers):
  :19   one more #include (the adapter header)
  :176  N_VNew_Parallel(...)                      -> N_VNew_B200(device context, ...)
  :242-257, :390  diffusion                        -> b200_diffusion_rhs          (5 ARKStep/LSRKStep constructors)
  :266, :394  (void*)&udata                       -> the B200 problem object built from udata
  :311  PSetup, PSolve                            -> b200_diffusion_psetup, b200_diffusion_psolve
  :352, :396  dom_eig                             -> b200_diffusion_domeig
Everything else -- option parsing, UserData::setup, Initial(), tolerances, methods, controllers, the evolve loop,
UserOutput, ARKodePrintAllStats -- is the reference's code, compiled as it is."""
import re
import sys

SUBS = [
    (r'#include "diffusion_2D\.hpp"\n', '#include "diffusion_2D.hpp"\n#include "refmain_adapter.hpp"\n', 1),
    (r'N_VNew_Parallel\(udata\.comm_c, udata\.nodes_loc, udata\.nodes, ctx\)',
     'N_VNew_B200(b200_adapter_ctx(), udata.nodes_loc, udata.nodes, ctx)', 1),
    (r'ARKStepCreate\(nullptr, diffusion,', 'ARKStepCreate(nullptr, b200_diffusion_rhs,', 1),
    (r'ARKStepCreate\(diffusion, nullptr,', 'ARKStepCreate(b200_diffusion_rhs, nullptr,', 1),
    (r'LSRKStepCreateSSP\(diffusion,', 'LSRKStepCreateSSP(b200_diffusion_rhs,', 1),
    (r'LSRKStepCreateSTS\(diffusion,', 'LSRKStepCreateSTS(b200_diffusion_rhs,', 2),
    (r'ARKodeSetUserData\((\w+), \(void\*\)&udata\)', r'ARKodeSetUserData(\1, b200_adapter_user_data(udata))', 2),
    (r'ARKodeSetPreconditioner\(arkode_mem, PSetup, PSolve\)',
     'ARKodeSetPreconditioner(arkode_mem, b200_diffusion_psetup, b200_diffusion_psolve)', 1),
    (r'LSRKStepSetDomEigFn\((\w+), dom_eig\)', r'LSRKStepSetDomEigFn(\1, b200_diffusion_domeig)', 2),
]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    text = open(src).read()
    for pat, rep, count in SUBS:
        text, n = re.subn(pat, rep, text)
        if n != count:
            raise SystemExit("patch_reference_main: pattern %r found %d times, expected %d" % (pat, n, count))
    open(dst, "w").write(text)


if __name__ == "__main__":
    main()
