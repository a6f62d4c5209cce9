"""Drop-in for the reference's ``pose_estimation_module`` (/root/reference/scripts/pose_estimation_module.py): same
function names, argument meaning and file formats, without the ROS ``tf`` / matplotlib imports at module load (the
reference imports both at the top, :8-11, and prints a banner, :6; ``tf`` is restated in transformations_lite and
matplotlib is imported lazily by the one plotting helper).  Host-side float64 helpers; none of this is a kernel.
"""
from __future__ import annotations

import numpy as np

from . import transformations_lite as _tf


def transformation_from_translation_quaternion(translation, quaternion):          # reference :15-23
    rotation_matrix = _tf.quaternion_matrix(quaternion)[:3, :3]
    transformation_matrix = np.eye(4)
    transformation_matrix[:3, :3] = rotation_matrix
    transformation_matrix[:3, 3] = translation
    return transformation_matrix


def translation_from_transformation_matrix(transformation_matrix):                # reference :26-28
    return [transformation_matrix[0, 3], transformation_matrix[1, 3], transformation_matrix[2, 3]]


def rotation_matrix_to_quaternion(rotation_matrix):                               # reference :31-57 (Shepperd), [x, y, z, w]
    m = rotation_matrix
    trace = np.trace(m)
    if trace > 0:
        S = np.sqrt(trace + 1.0) * 2.0
        w = 0.25 * S
        x = (m[2, 1] - m[1, 2]) / S
        y = (m[0, 2] - m[2, 0]) / S
        z = (m[1, 0] - m[0, 1]) / S
    elif (m[0, 0] > m[1, 1]) and (m[0, 0] > m[2, 2]):
        S = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2.0
        w = (m[2, 1] - m[1, 2]) / S
        x = 0.25 * S
        y = (m[0, 1] + m[1, 0]) / S
        z = (m[0, 2] + m[2, 0]) / S
    elif m[1, 1] > m[2, 2]:
        S = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2.0
        w = (m[0, 2] - m[2, 0]) / S
        x = (m[0, 1] + m[1, 0]) / S
        y = 0.25 * S
        z = (m[1, 2] + m[2, 1]) / S
    else:
        S = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2.0
        w = (m[1, 0] - m[0, 1]) / S
        x = (m[0, 2] + m[2, 0]) / S
        y = (m[1, 2] + m[2, 1]) / S
        z = 0.25 * S
    return [x, y, z, w]


def quaternion_from_transformation_matrix(transformation_matrix):                 # reference :60-65
    return rotation_matrix_to_quaternion(transformation_matrix[:3, :3])


def get_marker_to_marker_transformation(previous_cTm_transform, current_cTm_transform):   # reference :68-71
    return np.matmul(np.linalg.inv(previous_cTm_transform), current_cTm_transform)


def get_camera_to_camera_transformation(previous_cTm_transform, current_cTm_transform):   # reference :74-77
    return np.matmul(previous_cTm_transform, np.linalg.inv(current_cTm_transform))


def format_stamped_line(timestamp, translation, quaternion):
    """One line of a stamped trajectory file exactly as the reference writes it (:80-86): eight str() fields separated by
    single spaces, a trailing space, newline.  (Python 3 str(float) is the shortest round-trip repr; the reference ran
    under Python 2 whose str(float) is %.12g -- both parse identically with float()/np.genfromtxt, which is how every
    reader in the reference consumes these files.)"""
    return (str(timestamp) + " " + str(translation[0]) + " " + str(translation[1]) + " " + str(translation[2]) + " "
            + str(quaternion[0]) + " " + str(quaternion[1]) + " " + str(quaternion[2]) + " " + str(quaternion[3]) + " " + "\n")


def write_to_output_file(output_file_path, timestamp, translation, quaternion):   # reference :80-86
    with open(output_file_path, "a") as file:
        file.write(format_stamped_line(timestamp, translation, quaternion))


def clear_txt_file_contents(file_path):                                           # reference :89-91
    with open(file_path, "w") as file:
        file.truncate()


def get_velocity_between_timestamps(relative_position_change, previous_timestamp, current_timestamp):   # reference :94-111
    time_change = current_timestamp - previous_timestamp
    translation = np.array([relative_position_change[0, 3], relative_position_change[1, 3], relative_position_change[2, 3]])
    translation_velocity = translation / time_change
    rotational_velocity = relative_position_change[:3, :3] / time_change
    velocity_transformation = np.eye(4)
    velocity_transformation[:3, :3] = rotational_velocity
    velocity_transformation[:3, 3] = translation_velocity
    return velocity_transformation


def get_gt_vo_difference(gt_file_path, vo_file_path):                             # reference :113-127 (returns on the first row)
    ground_truth_data = np.genfromtxt(gt_file_path)
    vis_odom_data = np.genfromtxt(vo_file_path)
    for i in range(ground_truth_data.shape[0] - 1):
        gt_euler = np.array(_tf.euler_from_quaternion(tuple(ground_truth_data[i, 4:8])))
        vo_euler = np.array(_tf.euler_from_quaternion(tuple(vis_odom_data[i, 4:8])))
        return vo_euler - gt_euler


def write_gt_vo_difference_to_file(gt_file_path, vo_file_path, output_file_path):   # reference :130-147
    ground_truth_data = np.genfromtxt(gt_file_path)
    vis_odom_data = np.genfromtxt(vo_file_path)
    with open(output_file_path, "w") as file:
        for i in range(ground_truth_data.shape[0] - 1):
            timestamp = ground_truth_data[i, 0]
            gt_euler = np.array(_tf.euler_from_quaternion(tuple(ground_truth_data[i, 4:8])))
            vo_euler = np.array(_tf.euler_from_quaternion(tuple(vis_odom_data[i, 4:8])))
            gt_vo_difference = vo_euler - gt_euler
            print(gt_vo_difference)
            file.write("at timestamp {} the gt vo euler angle difference is {} \n".format(timestamp, gt_vo_difference))


def append_transformation_to_file(transformation_matrix, file_path):              # reference :150-154
    with open(file_path, "a") as file:
        for row in transformation_matrix:
            file.write(" ".join(str(value) for value in row) + "\n")


def compute_gt_vo_translation_difference(gt_file_path, vo_file_path):             # reference :156-164
    ground_truth = np.genfromtxt(gt_file_path)
    vis_odom = np.genfromtxt(vo_file_path)
    d = np.array(vis_odom[1:4]) - np.array(ground_truth[1:4])
    return [d[0], d[1], d[2]]


def visualize_gt_vo_translation_difference(translation_difference, plot_output_path):   # reference :168-184
    try:
        import matplotlib.pyplot as plt   # not installed in the build image; plotting is outside the hot path
    except ImportError as e:
        raise ImportError("visualize_gt_vo_translation_difference needs matplotlib") from e
    d = translation_difference
    fig = plt.figure(figsize=(5, 5))
    ax = fig.add_subplot(111, projection="3d")
    ax.plot([d[0], d[0]], [d[1], d[1]], [d[2], d[2]], "bo-")
    ax.scatter(d[0], d[1], d[2], c="r", marker="o", label="Vector 1")
    ax.scatter(d[0], d[1], d[2], c="g", marker="o", label="Vector 2")
    ax.set_xlabel("X"); ax.set_ylabel("Y"); ax.set_zlabel("Z")
    ax.set_title("Translation Difference Visualization")
    ax.legend()
    plt.savefig(plot_output_path, format="jpg", dpi=300)
