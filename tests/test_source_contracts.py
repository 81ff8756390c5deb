"""Static checks of contracts the kernel sources keep by construction (no GPU needed).

Programmatic dependent launch (csrc/common.cuh): a kernel launched with the programmatic-stream-serialisation attribute may start
before the kernel ahead of it on the stream has finished, so it MUST execute `griddepcontrol.wait` before it reads what that kernel
produced or writes anything.  A kernel that is launched through `launch_pdl` without containing the wait would pass most tests
(the race window is short) and read stale data once in a while — so the pairing is checked on the source text."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "deepfake_video_detection_b200", "csrc")


def _global_bodies(text):
    """(name, body) of every __global__ function of a translation unit (brace matching from the first '{' after the signature)."""
    out = []
    for m in re.finditer(r"__global__", text):
        sig_end = text.index("{", text.index("(", m.end()))
        # the signature's own parentheses may contain braces only in default arguments (none here): first '{' after ')' opens the body
        depth, i = 0, sig_end
        while True:
            c = text[i]
            if c == "{":
                depth += 1
            elif c == "}":
                depth -= 1
                if depth == 0:
                    break
            i += 1
        head = text[m.end():sig_end]
        name = re.findall(r"([A-Za-z_][A-Za-z_0-9]*)\s*\(", head)[-1] if "(" in head else "?"
        out.append((name, text[sig_end:i + 1]))
    return out


def test_every_programmatically_launched_kernel_waits_for_its_predecessor():
    pdl_files, plain_files = [], []
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith(".cu"):
            continue
        text = open(os.path.join(CSRC, f)).read()
        (pdl_files if "launch_pdl(" in text else plain_files).append((f, text))
    assert len(pdl_files) >= 7, "the EfficientNet step's kernels are launched through launch_pdl"
    for f, text in pdl_files:
        assert "<<<" not in text, f"{f}: mixes launch_pdl with plain launches — every kernel of the file must follow one contract"
        bodies = _global_bodies(text)
        assert bodies, f
        for name, body in bodies:
            assert "griddep_wait();" in body, f"{f}: kernel {name} is launched with the programmatic attribute but never waits"
    for f, text in plain_files:
        assert "griddep_wait" not in text, f"{f}: waits for a programmatic dependency but is launched plainly"


def test_wait_precedes_every_global_store_in_the_marching_kernels():
    """In the two marching kernels the wait must come before the first cp.async of activations (issue_row) and before any store."""
    for f in ("dwconv_march.cu", "mbconv_fused.cu"):
        text = open(os.path.join(CSRC, f)).read()
        (name, body), = _global_bodies(text)
        wait = body.index("griddep_wait();")
        first_issue = body.index("issue_row();")          # the first CALL (the lambda's definition is `auto issue_row = ...`)
        assert wait < first_issue, f"{f}: activations are fetched before the dependency on the previous kernel resolves"
        for store in ("orow + j * C", "partials +"):
            assert wait < body.index(store), f"{f}: {store!r} before the wait"
