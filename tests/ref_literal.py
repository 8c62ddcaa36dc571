"""Second, independent restatement of the reference's Rust hot path: a literal, loop-for-loop Python transcription of
rust/heimdall-core/src/detection.rs:127-317 written straight from the Rust source (NOT from oracle/hv_oracle.c).
Pure-Python loops, so only for small images; used to cross-check the C oracle (tests/test_oracle.py)."""
import math


def f64_as_u8(v):
    if v != v:
        return 0
    return max(0, min(255, int(v)))  # int() truncates toward zero


def f64_as_i32(v):
    if v != v:
        return 0
    if v >= 2147483647.0:
        return 2147483647
    if v <= -2147483648.0:
        return -2147483648
    return int(v)


def detect_contamination(image, min_size=10.0, max_size=3000.0, threshold=25.0):
    """image: nested list / ndarray [h][w][c].  Returns dict(gray, blurred, binary, labels, defects)."""
    height, width, channels = len(image), len(image[0]), len(image[0][0])
    gray = [[0] * width for _ in range(height)]
    if channels == 3:
        for i in range(height):
            for j in range(width):
                r, g, b = int(image[i][j][0]), int(image[i][j][1]), int(image[i][j][2])
                gray[i][j] = f64_as_u8(0.299 * float(r) + 0.587 * float(g) + 0.114 * float(b))
    elif channels == 1:
        for i in range(height):
            for j in range(width):
                gray[i][j] = int(image[i][j][0])
    else:
        raise ValueError("Invalid image dimensions: expected 3D array")

    blur_radius = 2
    blurred = [row[:] for row in gray]
    for i in range(blur_radius, height - blur_radius):
        for j in range(blur_radius, width - blur_radius):
            s = 0
            count = 0
            for bi in range(-blur_radius, blur_radius + 1):
                for bj in range(-blur_radius, blur_radius + 1):
                    s += gray[i + bi][j + bj]
                    count += 1
            blurred[i][j] = (s // count) & 0xFF

    window_size = 11
    c = f64_as_i32(threshold)
    binary = [[0] * width for _ in range(height)]
    for i in range(height):
        for j in range(width):
            start_i = max(i - window_size // 2, 0)
            end_i = min(i + window_size // 2, height - 1)
            start_j = max(j - window_size // 2, 0)
            end_j = min(j + window_size // 2, width - 1)
            s = 0
            count = 0
            for y in range(start_i, end_i + 1):
                for x in range(start_j, end_j + 1):
                    s += blurred[y][x]
                    count += 1
            mean = s // count
            rhs = (mean - c + 2 ** 31) % 2 ** 32 - 2 ** 31  # i32 arithmetic of a release build wraps (no overflow-checks)
            binary[i][j] = 255 if blurred[i][j] < rhs else 0

    defects = []
    labels = [[0] * width for _ in range(height)]
    visited = [[False] * width for _ in range(height)]
    ncomp = 0
    for i in range(height):
        for j in range(width):
            if binary[i][j] == 255 and not visited[i][j]:
                ncomp += 1
                pixels = []
                stack = [(i, j)]
                visited[i][j] = True
                while stack:
                    y, x = stack.pop()
                    pixels.append((y, x))
                    labels[y][x] = ncomp
                    for ny, nx in ((max(y - 1, 0), x), (y + 1, x), (y, max(x - 1, 0)), (y, x + 1)):
                        if ny < height and nx < width and binary[ny][nx] == 255 and not visited[ny][nx]:
                            stack.append((ny, nx))
                            visited[ny][nx] = True
                area = float(len(pixels))
                if area >= min_size and area <= max_size:
                    center_y = sum(p[0] for p in pixels) // len(pixels)
                    center_x = sum(p[1] for p in pixels) // len(pixels)
                    fg_sum = bg_sum = fg_count = bg_count = 0
                    margin = 2
                    for y in range(max(center_y - margin, 0), min(center_y + margin, height - 1) + 1):
                        for x in range(max(center_x - margin, 0), min(center_x + margin, width - 1) + 1):
                            if binary[y][x] == 255:
                                fg_sum += gray[y][x]
                                fg_count += 1
                            else:
                                bg_sum += gray[y][x]
                                bg_count += 1
                    fg_mean = fg_sum / fg_count if fg_count > 0 else 127.0
                    bg_mean = bg_sum / bg_count if bg_count > 0 else 127.0
                    intensity_diff = abs(bg_mean - fg_mean)
                    ys = [p[0] for p in pixels]
                    xs = [p[1] for p in pixels]
                    rect_area = (max(ys) - min(ys) + 1) * (max(xs) - min(xs) + 1)
                    shape_score = 1.0 - (area / float(rect_area)) if rect_area > 0 else 0.5
                    intensity_score = min(intensity_diff / 30.0, 1.0)
                    confidence = (intensity_score * 0.7) + (shape_score * 0.3)
                    if confidence >= 0.3:
                        defects.append({"position": (center_y, center_x), "size": area, "confidence": confidence,
                                        "label": ncomp})
    return {"gray": gray, "blurred": blurred, "binary": binary, "labels": labels, "ncomp": ncomp, "defects": defects}
