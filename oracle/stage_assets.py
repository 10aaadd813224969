"""TEST INFRASTRUCTURE: stages the tinyapp's assets from the reference tree into oracle/_ref/assets (git-ignored; travels to
the GPU box with the prebuilt checkers because /root/reference does not exist there). One texture the glTF scene names,
pica/textures/Wax_Pastel_Label_02_baseColor.png, is missing from the reference snapshot; tinygltf then leaves the
image empty and HostScene::AddScene (host_scene.cpp:313-320) reads through a null pointer. A 64x64 flat pastel PNG is written
in its place so that the scene loads; everything else is copied byte for byte.

usage: python stage_assets.py <reference root> <destination dir>"""
import os
import shutil
import struct
import sys
import zlib


def write_png(path, w, h, rgba):
    raw = b"".join(b"\x00" + bytes(rgba) * w for _ in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b""))


def main():
    ref, dst = sys.argv[1], sys.argv[2]
    shared = os.path.join(ref, "apps", "_shareddata")
    os.makedirs(dst, exist_ok=True)
    shutil.copytree(os.path.join(shared, "pica"), os.path.join(dst, "pica"), dirs_exist_ok=True)
    for name in ("legocar.obj", "legocar.mtl", "CesiumMan.glb"):      # CesiumMan: the skinned, animated scene imguiapp / viewerapp use
        shutil.copy(os.path.join(shared, name), os.path.join(dst, name))
    shutil.copy(os.path.join(ref, "apps", "tinyapp", "camera.xml"), os.path.join(dst, "camera.xml"))
    # legocar.mtl names "textures/legoshld.tga" (resolved against the working directory, case-insensitively on the reference's
    # platform); the file in the tree is textures/LEGOSHLD.tga. The host runs with the staged directory as working directory.
    os.makedirs(os.path.join(dst, "textures"), exist_ok=True)
    shutil.copy(os.path.join(shared, "textures", "LEGOSHLD.tga"), os.path.join(dst, "textures", "legoshld.tga"))
    for root, dirs, files in os.walk(dst):
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
    missing = os.path.join(dst, "pica", "textures", "Wax_Pastel_Label_02_baseColor.png")
    if not os.path.exists(missing):
        write_png(missing, 64, 64, (214, 196, 170, 255))
    open(os.path.join(dst, ".staged"), "w").close()


if __name__ == "__main__":
    main()
