"""Extracts the known-answer vectors of the reference's own unit test
/root/reference/src/c++/lib/starling_common/test/starling_read_align_test.cpp (test_make_start_pos_alignment :67-170,
test_end_pin_start_pos :200-355) into tests/golden/read_align_unit_goldens.json.  Run in the build container (the GPU box has no
/root/reference); the JSON is committed.  Both test helpers place ONE indel next to the fixed 1 bp deletion at 1075 on a 100 bp
read: make_start_pos_alignment from (ref 1000, read_start), get_end_pin_start_pos from (ref end 1100, read_end)."""
import json
import os
import re

SRC = "/root/reference/src/c++/lib/starling_common/test/starling_read_align_test.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "read_align_unit_goldens.json")


def blocks(text):
    depth, start = 0, None
    for i, c in enumerate(text):
        if c == "{":
            if depth == 0:
                start = i
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                yield text[start : i + 1]


def key_of(block, ins10):
    m = re.search(r"IndelKey ik\((\d+),\s*INDEL::INDEL(?:,\s*(\d+))?(?:,\s*(\w+))?\)", block)
    if not m:
        return None
    return {"pos": int(m.group(1)), "del_len": int(m.group(2) or 0), "ins": ins10 if m.group(3) else ""}


def main():
    text = open(SRC).read()
    ins10 = re.search(r'insertSeq10 = "(\w+)"', text).group(1)
    a = text.index("BOOST_AUTO_TEST_CASE( test_make_start_pos_alignment )")
    b = text.index("test_end_pin_indel_placement(")
    c = text.index("BOOST_AUTO_TEST_CASE( test_end_pin_start_pos )")
    d = text.index("BOOST_AUTO_TEST_CASE( test_realign_and_score_read )")
    start_cases, end_cases = [], []
    body = text[a:b]
    for blk in blocks(body[body.index("{") + 1 :]):
        k = key_of(blk, ins10)
        if k is None:
            continue
        rs = re.search(r"test_indel_placement\(ik(?:,(\d+))?\)", blk)
        case = {"key": k, "read_start": int(rs.group(1) or 0), "path": re.search(r'path_compare\("([^"]+)"', blk).group(1)}
        for side in ("leading", "trailing"):
            if re.search(rf"cal\.{side}_indel_key,ik\)", blk):
                case[side] = True
            elif re.search(rf"cal\.{side}_indel_key\.type,INDEL::NONE", blk):
                case[side] = False
        m = re.search(r"cal\.al\.pos,(\d+)", blk)
        if m:
            case["pos"] = int(m.group(1))
        start_cases.append(case)
    body = text[c:d]
    for blk in blocks(body[body.index("{") + 1 :]):
        k = key_of(blk, ins10)
        if k is None:
            continue
        re_ = re.search(r"test_end_pin_indel_placement\(ik(?:,(\d+))?\)", blk)
        case = {"key": k, "read_end": int(re_.group(1) or 100)}
        if "BOOST_CHECK_THROW" in blk:
            case["throws"] = True
        else:
            case["ref_start"] = int(re.search(r"res\.first,(\d+)", blk).group(1))
            case["read_start"] = int(re.search(r"res\.second,(\d+)", blk).group(1))
        end_cases.append(case)
    json.dump({"source": "starling_common/test/starling_read_align_test.cpp", "fixed_key": {"pos": 1075, "del_len": 1, "ins": ""}, "read_length": 100,
               "ref_start": 1000, "ref_end": 1100, "make_start_pos_alignment": start_cases, "get_end_pin_start_pos": end_cases}, open(OUT, "w"), indent=1)
    print(len(start_cases), "make_start_pos_alignment cases,", len(end_cases), "get_end_pin_start_pos cases ->", OUT)


if __name__ == "__main__":
    main()
